#!/usr/bin/env python
"""Warp-stall samples of one kernel from an ncu report, aggregated per ROLE (ranges of source lines) and stall reason.

   python scripts/ncu_roles.py <prof.ncu-rep> <kernel regex> <lib.so> <cu file stem> name:first:last [name:first:last ...]
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kre, lib, stem = sys.argv[1:5]
roles = [(n, int(a), int(b)) for n, a, b in (x.split(':') for x in sys.argv[5:])]
r = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kre}'], capture_output=True, text=True)
rows = list(csv.reader(r.stdout.splitlines()))
hi = next(i for i, x in enumerate(rows) if x and x[0] == 'Address')
hdr = rows[hi]
data = []
for x in rows[hi + 1:]:
    if x and x[0] == 'Kernel Name':
        break
    if len(x) >= len(hdr):
        data.append(x)
si = hdr.index('# Samples')
ei = hdr.index('Warp Instructions Executed') if 'Warp Instructions Executed' in hdr else (hdr.index('# Instructions Executed') if '# Instructions Executed' in hdr else None)
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith(stem + '.')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
linemap, cur_main, in_fn = {}, None, False
for ln in dis:
    if ln.startswith('.text.') and ln.rstrip().endswith(':'):
        in_fn = re.search(kre, ln) is not None
        cur_main = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if os.path.basename(m.group(1)).startswith(stem + '.'):
            cur_main = int(m.group(2))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m and in_fn:
        linemap[int(m.group(1), 16) // 16] = cur_main
tot = sum(int(x[si] or 0) for x in data)
agg = collections.defaultdict(collections.Counter)
inst = collections.Counter()
ninst = collections.Counter()
for i, x in enumerate(data):
    line = linemap.get(i) or -1
    role = next((n for n, a, b in roles if a <= line <= b), 'other')
    ninst[role] += 1
    if ei is not None and x[ei]:
        inst[role] += int(float(x[ei]))
    for j, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h and x[j] not in ('', '0'):
            agg[role][h[6:]] += int(x[j])
    agg[role]['TOTAL'] += int(x[si] or 0)
print(f'kernel {kre}: {tot} samples, {len(data)} SASS instructions; columns: {[h for h in hdr if "nstr" in h]}')
for n, _, _ in roles + [('other', 0, 0)]:
    c = agg[n]
    t = c.pop('TOTAL', 0)
    print(f'{n:10s} {t:7d} samples {100 * t / max(tot, 1):5.1f}%  sass {ninst[n]:5d}  executed {inst[n]:10d} | ' + ' '.join(f'{k}:{100 * v / max(t, 1):.0f}%' for k, v in c.most_common(7)))

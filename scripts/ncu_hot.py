#!/usr/bin/env python
"""Hot spots of one kernel from an ncu report: warp-stall samples per SASS instruction (source page), mapped to CUDA
source lines through nvdisasm -g of the cubin inside the shared library, and aggregated per source line.

   python scripts/ncu_hot.py <prof.ncu-rep> <kernel regex> <lib.so> <cu file stem> [top N]
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kre, lib, stem = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
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
# line map
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith(stem + '.')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
linemap, cur_main, cur, in_fn = {}, None, None, False
for ln in dis:
    if ln.startswith('.text.') and ln.rstrip().endswith(':'):
        in_fn = re.search(kre, ln) is not None
        cur_main = cur = None
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        if os.path.basename(m.group(1)).startswith(stem + '.'):
            cur_main = cur
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', ln)
    if m and in_fn:
        linemap[int(m.group(1), 16) // 16] = (cur_main, cur)
tot = sum(int(x[si] or 0) for x in data)
per_line = collections.Counter()
reasons_line = collections.defaultdict(collections.Counter)
for i, x in enumerate(data):
    n = int(x[si] or 0)
    main, inner = linemap.get(i, (None, None))
    key = main[1] if main else -1
    per_line[key] += n
    for j, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h and x[j] not in ('', '0'):
            reasons_line[key][h[6:]] += int(x[j])
src = open(os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc', 'tc', stem + '.cu')).read().splitlines() if os.path.exists(
    os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc', 'tc', stem + '.cu')) else []
print(f'kernel {kre}: {tot} samples over {len(data)} SASS instructions')
for line, n in per_line.most_common(top):
    text = src[line - 1].strip()[:95] if 0 < line <= len(src) else '?'
    rs = ' '.join(f'{k}:{v}' for k, v in reasons_line[line].most_common(3))
    print(f'{n:7d} {100*n/tot:5.1f}%  L{line:<4d} {text:95s} | {rs}')

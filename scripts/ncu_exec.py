#!/usr/bin/env python
"""Warp-level instructions executed per CUDA source line of one kernel, from an ncu report (source page).
   python scripts/ncu_exec.py <prof.ncu-rep> <kernel regex> <lib.so> <cu file stem> [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kre, lib, stem = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
r = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kre}'], capture_output=True, text=True)
rows = list(csv.reader(r.stdout.splitlines()))
hi = next(i for i, x in enumerate(rows) if x and x[0] == 'Address')
hdr = rows[hi]
data = [x for x in rows[hi + 1:] if len(x) >= len(hdr)]
ie, isrc = hdr.index('Instructions Executed'), hdr.index('Source')
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
per, spin = collections.Counter(), 0
for i, x in enumerate(data):
    n = int(x[ie] or 0)
    per[linemap.get(i, -1)] += n
tot = sum(per.values())
src = open(os.path.join(os.path.dirname(os.path.abspath(lib)), 'csrc', 'tc', stem + '.cu')).read().splitlines()
print(f'{tot} warp instructions')
for line, n in per.most_common(top):
    text = src[line - 1].strip()[:110] if line and 0 < line <= len(src) else '?'
    print(f'{n:12d} {100*n/tot:5.1f}%  L{line}  {text}')

#!/usr/bin/env python
"""Count the tensor-core / TMEM / async-copy SASS instructions per kernel of libminer_b200.so (cuobjdump -sass; no GPU needed).

    python scripts/sass_evidence.py [lib.so] > profiles/rNN_sass_evidence.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'miner_b200', 'libminer_b200.so')
KEYS = [('UTCHMMA', r'\bUTCHMMA'), ('UTCBAR', r'\bUTCBAR'), ('LDTM', r'\bLDTM'), ('STTM', r'\bSTTM'), ('UTMALDG', r'\bUTMALDG'),
        ('LDGSTS', r'\bLDGSTS'), ('SYNCS', r'\bSYNCS'), ('REDUX', r'\bREDUX'), ('MUFU.TANH', r'MUFU\.TANH'), ('MUFU.EX2', r'MUFU\.EX2'),
        ('LDG.128', r'LDG\.E\.(?:\w+\.)*128'), ('LDS.128', r'LDS\.128'), ('FFMA2', r'\bFFMA2'), ('DFMA', r'\bDFMA')]
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
counts, order, total = {}, [], {}
cur = None
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        order.append(cur)
        continue
    if cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        total[cur] += 1
        for k, pat in KEYS:
            if re.search(pat, line):
                counts[cur][k] += 1
print('SASS evidence (cuobjdump -sass of the sm_100a cubins inside %s): instructions per kernel.' % os.path.relpath(lib, ROOT))
print('UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, LDGSTS = cp.async,')
print('SYNCS = mbarrier ops, REDUX = redux.sync, FFMA2 = packed f32x2 FMA; `n` = SASS instructions of the kernel.\n')
def demangle(n):
    r = subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    r = re.sub(r'\(anonymous namespace\)::', '', r)
    return r.split('(')[0].replace('void ', '')[:64]
for name in sorted(order, key=demangle):
    c = counts[name]
    print('%-64s n:%-6d %s' % (demangle(name), total[name], ' '.join('%s:%d' % (k, c[k]) for k, _ in KEYS if c[k])))

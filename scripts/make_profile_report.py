#!/usr/bin/env python
"""Turn the ncu outputs of scripts/gpu_profile.sh (gpurun_out/launches.csv, gpurun_out/prof.ncu-rep) into the tracked
evidence under profiles/:  <tag>_ncu_summary.txt (launch list + full-set metrics + hot source lines) and traffic.json
(DRAM bytes per impression of each fused kernel, read by bench.py for roofline.traffic).

    python scripts/make_profile_report.py r01_v4 "<what was profiled>" [impressions_per_launch]
"""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, title = sys.argv[1], sys.argv[2]
per_launch = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
rep = os.path.join(ROOT, 'gpurun_out', 'prof.ncu-rep')
out = os.path.join(ROOT, 'profiles', f'{tag}_ncu_summary.txt')
subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), os.path.join(ROOT, 'gpurun_out', 'launches.csv'), rep, out, title],
               stdout=subprocess.DEVNULL, check=True)
lib = os.path.join(ROOT, 'miner_b200', 'libminer_b200.so')
with open(out, 'a') as f:
    for kre, stem in (('tscore_x_kernel', 'tscore_x_kernel'), ('tscore_kernel', 'tscore_kernel'), ('hist_kernel2', 'hist_kernel2'), ('cand_kernel', 'cand_kernel')):
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_hot.py'), rep, kre, lib, stem, '14'], capture_output=True, text=True)
        if r.returncode == 0 and r.stdout.strip():
            f.write(f'\n== hot source lines, warp-stall samples ({kre}) ==\n' + r.stdout)
# traffic
r = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True)
rows = list(csv.reader(r.stdout.splitlines()))
hdr, units = rows[0], rows[1]
tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
old = json.load(open(tpath)) if os.path.exists(tpath) else {}
traffic = {}
for row in rows[2:]:
    name = row[hdr.index('Kernel Name')]
    key = 'tscore_kernel' if ('tscore_kernel' in name or 'tscore_x_kernel' in name) else 'hist_kernel' if 'hist_kernel' in name else 'cand_kernel' if 'cand_kernel' in name else None
    if not key:
        continue
    tot = 0.0
    for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        i = hdr.index(m)
        v = float(row[i].replace(',', ''))
        tot += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
    t = traffic.setdefault(key, {'samples': []})
    t['samples'].append(tot / per_launch)
for k, t in traffic.items():
    t['dram_bytes_per_impression'] = sum(t['samples']) / len(t['samples'])
    t['source'] = f'profiles/{tag}_ncu_summary.txt (ncu --set full, {per_launch} impressions per launch)'
    del t['samples']
old.update(traffic)          # kernels not in this capture keep their earlier figures
traffic = old
json.dump(traffic, open(tpath, 'w'), indent=1)
print(open(out).read()[:3000])
print(json.dumps(traffic, indent=1))

#!/bin/bash
# One box session: the metric / pipeline tests that changed, rank_metrics A/B (resident blocks of the lane kernel), the stage-by-stage
# timing of a device step, then a bench line.  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -q -x --timeout 120 -k "rank_metrics or host_evaluator or evaluators or properties_at_scale" > gpurun_out/t_rm.log 2>&1
echo "t_rm exit $?" >> gpurun_out/t_rm.log
tail -4 gpurun_out/t_rm.log
for lib in libminer_b200.so libminer_b200_rm3.so; do
  echo "== $lib"
  MINER_B200_LIB=miner_b200/$lib timeout 120 python scripts/time_metrics.py 2>&1 | tail -1
done | tee gpurun_out/rm_ab.txt
timeout 200 python scripts/step_gap.py 1000000 6 2>&1 | tee gpurun_out/step_gap.txt
timeout -k 10 400 python bench.py --steps 3 --warmup 3 --cpu-sample 1000 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print([(k['kernel'][:20], round(k['ms_per_step'], 3)) for k in d['kernels']])
print('parity', d['parity']['order_flips'], d['parity']['scores_normwise'], d['parity']['metrics_abs_diff'])
PY

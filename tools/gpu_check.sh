#!/bin/bash
# usage (on the GPU box via gpurun): tools/gpu_check.sh [tests] [ncu]
# runs the gpu tests (optional), a short bench, and (optional) an ncu --set full capture of the fused kernel
set -o pipefail
mkdir -p gpurun_out
if [[ "$*" == *tests* ]]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -6; fi
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err || { tail -20 gpurun_out/bench.err; exit 1; }
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("ms/step %.4f  value %.4g  kernel_ms %.4f  step_frac %.4f  e2e_ms %.4f launches %d" % (d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["config"]["step_hbm_frac"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
PY
if [[ "$*" == *ncu* ]]; then
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_project_weight -s 4 -c 1 -f -o gpurun_out/prof_weight python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_w.log 2>&1
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 33 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
fi

#!/bin/bash
# usage (on the GPU box via gpurun): tools/gpu_check.sh [tests] [ncu] [sweep]
# runs the gpu tests (optional), a short bench, and (optional) ncu captures of the weight kernels + the launch list
set -o pipefail
mkdir -p gpurun_out
if [[ "$*" == *tests* ]]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
summ() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.4f  value %.4g  kernel_ms %.4f  step_frac %.4f  e2e_ms %.4f launches %d deferred %s" % (d["ms_per_step"], d["value"], d["roofline"]["kernel_ms"], d["config"]["step_hbm_frac"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["config"].get("generic_path_particles")))
PY
}
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err || { tail -20 gpurun_out/bench.err; exit 1; }
summ gpurun_out/bench.json
if [[ "$*" == *sweep* ]]; then
  ESLAM_GPU_NO_FAST=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nofast.json 2>> gpurun_out/bench.err && { echo "generic only"; summ gpurun_out/bench_nofast.json; }
fi
if [[ "$*" == *ncu* ]]; then
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_weight_lean -s 4 -c 1 -f -o gpurun_out/prof_lean python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_w.log 2>&1
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 45 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
fi

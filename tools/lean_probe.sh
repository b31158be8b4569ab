#!/bin/bash
# usage (GPU box): tools/lean_probe.sh  -- step time and cheap ncu counters of the lean kernel for both CTA sizes
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.avg
for t in 1024 512; do
  export ESLAM_GPU_LEAN_THREADS=$t
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_t$t.json 2>> gpurun_out/bench.err
  python - gpurun_out/bench_t$t.json <<'PY'
import json,sys,os
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("threads %s: ms/step %.4f kernel_ms %.4f deferred %s" % (os.environ["ESLAM_GPU_LEAN_THREADS"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"].get("generic_path_particles")))
PY
  ncu --metrics $M --clock-control none -k regex:"k_weight_lean|k_project_bin|k_bin_scatter" -s 9 -c 3 --csv --log-file gpurun_out/probe_t$t.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - gpurun_out/probe_t$t.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
for r in rows[hdr+1:]:
    print("   %-28s %-70s %s" % (r[4][:28], r[-3], r[-1]))
PY
done

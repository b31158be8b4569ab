#!/bin/bash
# usage (GPU box, one GPU): tools/profile_r02.sh -- the round's ncu evidence: launch list of the bench command and --set full
# captures of the kernels this round touched.  Every ncu run is preceded by the same command without ncu (it must exit 0).
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 54 -c 27 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_l.log 2>&1
for k in k_weight_lean k_resample_stream k_norm_scan k_chain k_cdf k_tile_part k_loop_b k_project_bin; do
  ncu --set full --clock-control none --import-source on -k regex:"^$k" -s 4 -c 1 -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/raw_$k.csv 2>/dev/null
  python tools/ncu_raw_summary.py gpurun_out/raw_$k.csv "$k, ncu --set full --clock-control none ($B, C3)" > gpurun_out/${k}_ncu_raw.txt 2>/dev/null
done
ncu -i gpurun_out/prof_k_weight_lean.ncu-rep --page source --csv > gpurun_out/src_k_weight_lean.csv 2>/dev/null
rm -f gpurun_out/prof_k_norm_scan.ncu-rep gpurun_out/prof_k_chain.ncu-rep gpurun_out/prof_k_cdf.ncu-rep gpurun_out/prof_k_tile_part.ncu-rep gpurun_out/prof_k_loop_b.ncu-rep gpurun_out/prof_k_project_bin.ncu-rep
ls -la gpurun_out | head -40

#!/bin/bash
# ncu evidence for the DEFAULT path (consistency shortcut) of one (1024, 683) share-matrix pass:
#  1. plain run (must exit 0 first), 2. launch list with per-launch durations, 3. one --set full capture of the
#  dominant kernels (decode, coefficient check, interpolation, difference check) exported as raw CSV.
mkdir -p gpurun_out
timeout 100 python tools/prof_share.py 1024 683 1024 > gpurun_out/prof_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/prof_plain.log; exit 1; }
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_default.csv \
  python tools/prof_share.py 1024 683 1024 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 150 ncu --set full --clock-control none -k 'regex:k_decompress_vv|k_fd_coefcheck|k_fd_interp|k_fd_polycheck|k_fd_share_limbs' -c 5 \
  --csv --page raw --log-file gpurun_out/r1_default_path_full.csv python tools/prof_share.py 1024 683 1024 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out

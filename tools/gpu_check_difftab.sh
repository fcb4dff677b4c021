#!/bin/bash
# GPU check of the fused difference-table kernel: share-path parity tests, then the launch list of one default-path step
mkdir -p gpurun_out
timeout 130 python -m pytest tests/test_gpu_share.py -x -q -m gpu -k "not sparse_items" > gpurun_out/t_share.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/t_share.log
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_default_v2.csv \
  python tools/prof_share.py 1024 683 1024 > gpurun_out/ncu_launches_v2.log 2>&1
echo "launch list rc=$?"
grep -E "k_fd_|k_decompress" gpurun_out/r1_launches_default_v2.csv | awk -F'","' '{print $5, $NF}' | cut -c1-40,200-

#!/bin/bash
# GPU check of the decode-free consistency shortcut: share-path parity tests, launch list of one default-path step, short bench
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_gpu_share.py -x -q -m gpu -k "not sparse_items" > gpurun_out/t_share.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -6 gpurun_out/t_share.log
DKGV_FD_BYTES=1 timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_default_v3.csv \
  python tools/prof_share.py 1024 683 1024 > gpurun_out/ncu_launches_v3.log 2>&1
echo "launch list rc=$?"
grep -E "k_fd_|k_decompress" gpurun_out/r1_launches_default_v3.csv | awk -F'","' '{print substr($5,1,18), $NF}'
[ $rc -eq 0 ] || exit 1
DKGV_FD_BYTES=1 timeout 150 python bench.py --no-cpu --no-finalization --steps 5 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_v3.err; cut -c1-400 gpurun_out/bench_v3.json

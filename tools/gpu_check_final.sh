#!/bin/bash
# final r1 check with the decode-free shortcut as the DEFAULT: share-path parity tests, smoke, one --set full capture of the
# dominant default-path kernels
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_share.py -x -q -m gpu -k "not sparse_items" > gpurun_out/t_share_default_on.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/t_share_default_on.log
timeout 30 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 60 ncu --set full --clock-control none -k 'regex:k_fd_coefpoint|k_fd_coefsign|k_fd_difftab' -c 3 \
  --csv --page raw --log-file gpurun_out/r1_default_path_v3_full.csv python tools/prof_share.py 1024 683 1024 > gpurun_out/ncu_full_v3.log 2>&1
echo "full capture rc=$?"

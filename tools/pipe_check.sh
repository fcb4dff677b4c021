#!/bin/bash
# pipelined headline: N = 1 and one rank of N = 8 / N = 4 / N = 2 on ONE GPU (--emulate-world), lanes sweep
python -m pytest tests/test_gpu_share.py -x -q -k "pipelined or sharded or asynchronous or share_the_fixed" > gpurun_out/pipe_test.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pipe_test.log
for cfg in "1 2" "1 3" "1 4" "8 2" "8 3" "8 4" "8 6" "4 3" "4 4" "2 3" "2 4"; do
  set -- $cfg
  python bench.py --quick --no-cpu --no-peak --emulate-world $1 --lanes $2 > gpurun_out/pipe_w$1_l$2.json 2> gpurun_out/pipe_w$1_l$2.err; echo "w$1 l$2 rc=$?"
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/pipe_w*_l*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f,"pipe ms",round(d["ms_per_step"],4),"repeats",[round(x,4) for x in d["pipeline_repeats_ms_per_step"]],"sync ms",round(d["sync_call"]["ms_per_step"],4),"bad",d["parity"]["bad_verdict_bits_pipelined"])
    except Exception as e: print(f,"ERR",e)
P

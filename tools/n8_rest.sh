#!/bin/bash
# the "rest of the step" at N = 8 under four conditions (one gpurun --gpus 8 call): sampler nvidia-smi / NVML / off, and off without the L2 flush
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --quick --no-cpu --no-peak > gpurun_out/n8_$name.json 2> gpurun_out/n8_$name.err
  echo "$name rc=$?"
}
run smi DKGV_BENCH_SAMPLER=smi
run nvml DKGV_BENCH_SAMPLER=nvml
run off DKGV_BENCH_SAMPLER=off
run off_noflush DKGV_BENCH_SAMPLER=off DKGV_BENCH_FLUSH=0
python - <<'P'
import json
for n in ("smi","nvml","off","off_noflush"):
    try:
        d=json.loads(open(f"gpurun_out/n8_{n}.json").read().strip().splitlines()[-1])
        ks=[round(k["kernel_ms"],3) for k in d["roofline"]["shortcut_kernels"]]
        print(n, round(d["ms_per_step"],3), round(d["value"]/1e6,1), ks, d["clocks"])
    except Exception as e: print(n,"ERR",e)
P

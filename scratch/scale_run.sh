#!/bin/bash
# strong-scaling run of bench.py at 1/2/4/8 GPUs of one box -> gpurun_out/$1_bench_n{N}.json
tag=${1:-scale}
for N in 1 2 4 8; do
  if [ $N = 1 ]; then timeout 300 python bench.py --no-extras > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; fi
done
python - <<PY
import json
for N in (1,2,4,8):
    try:
        txt=open("gpurun_out/${tag}_bench_n%d.json"%N).read()
        d=json.loads([l for l in txt.splitlines() if l.startswith("{")][0])
        print(N, "%.3e"%d["value"], "%.3f"%d["ms_per_step"], {k:round(v["ms"],3) for k,v in d["kernels"].items()}, "e2e %.3e"%d["e2e"]["value"], d["details"]["final_ess"])
    except Exception as e:
        print(N, "failed", e)
PY

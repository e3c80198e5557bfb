#!/bin/bash
# 8-GPU evidence (charged 8x: keep it short): peer-gather check on all ranks, then the default bench line with
# each exchange (weak scaling of C2 + the strong-scaling block of C4).
N=${1:-8}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/gpu/peer_gather_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
for g in peer nccl; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --gather $g > gpurun_out/multi_${N}_$g.json 2> gpurun_out/multi_${N}_$g.err
  tail -1 gpurun_out/multi_${N}_$g.json | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.4g ms %.4f e2e %.4g strong %s" % (d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value", 0), d.get("strong")))' $g
done

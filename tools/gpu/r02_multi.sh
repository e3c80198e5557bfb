#!/bin/bash
# multi-GPU evidence: peer-gather check, weak scaling (default line) + strong scaling block, both exchanges
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/gpu/peer_gather_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -25
for g in peer nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --gather $g > gpurun_out/multi_${N}_$g.json 2> gpurun_out/multi_${N}_$g.err
  tail -1 gpurun_out/multi_${N}_$g.json | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.4g ms %.4f e2e %.4g strong %s" % (d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value", 0), d.get("strong")))' $g
done
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/multi_1.json 2>/dev/null; tail -1 gpurun_out/multi_1.json | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("1gpu value %.4g ms %.4f e2e %.4g strong %s parity %s" % (d["value"], d["ms_per_step"], (d["e2e"] or {}).get("value", 0), d.get("strong"), d.get("parity")))'

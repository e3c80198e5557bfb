#!/bin/bash
# 2-GPU check of the exchange: in-sweep stores (default) and the separate pack + store kernel, then the bench lines.
# Every multi-rank command runs under its own timeout (a protocol bug must not eat the GPU budget).
mkdir -p gpurun_out
for m in 1 0; do
  echo "== AB2_PEER_IN_SWEEP=$m"
  AB2_PEER_IN_SWEEP=$m timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/gpu/peer_gather_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -2
done
for g in peer nccl; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-e2e --gather $g > gpurun_out/multi_2_$g.json 2> gpurun_out/multi_2_$g.err
  echo "$g exit $?"
  tail -1 gpurun_out/multi_2_$g.json | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print(sys.argv[1], "value %.4g ms %.4f strong ms %s launches %s" % (d["value"], d["ms_per_step"], (d.get("strong") or {}).get("ms_per_step"), d.get("gpu_launches")))' $g
done
AB2_PEER_IN_SWEEP=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-e2e --gather peer 2>/dev/null | tail -1 | python -c 'import sys,json
d=json.loads(sys.stdin.read()); print("peer(separate kernel)", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]))'

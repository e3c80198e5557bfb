#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/gpu/peer_gather_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -25

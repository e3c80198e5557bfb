"""torchrun script (>= 2 GPUs): the fused pack + NVLink peer-memory all-gather of the first-step policy
(ab2_gar_policy_allgather) against ncclAllGather of the packed policy, over several steps with changing
problems (exercises the double buffering and the ack flow control)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
import bench  # noqa: E402
import aligator_b200.gar as gar  # noqa: E402
from aligator_b200 import sharding  # noqa: E402

nx, nu, N, B = 12, 6, 20, 96
s = gar.CudaRiccatiBatch(nx, nu, 0, 0, nx, N, B, device=local)
s.peer_gather_setup(dist, rank, world)
stream = torch.cuda.current_stream().cuda_stream
ok = True
for step in range(1, 8):
    prob = bench.synth_batch_torch(torch, B, N, nx, nu, dev, 1000 * step + rank)
    s.set_problem(*prob, memspace=gar.AB2_DEVICE, stream=stream)
    s.sweep(1e-9, stream=stream)
    s.policy_allgather(stream=stream)
    s.policy_allgather_wait(stream=stream)
    pol = torch.empty(B, nu, nx + 1, dtype=torch.float64, device=dev)
    s.first_step_policy_into(pol, stream=stream)
    ref = sharding.all_gather_policy(torch, dist, pol, world)
    torch.cuda.synchronize()
    ptr, st = s.peer_gather_buffer()
    assert st == step
    class _View:  # zero-copy torch view of the library-owned receive buffer
        __cuda_array_interface__ = {"shape": (world * B, nu, nx + 1), "typestr": "<f8", "data": (ptr, False), "version": 3}
    got = torch.as_tensor(_View(), device=dev).clone()
    torch.cuda.synchronize()
    same = bool(torch.equal(got, ref))
    ok = ok and same
    if rank == 0:
        print("step", step, "peer-gathered == nccl-gathered:", same, flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEER_GATHER_OK" if int(flag.item()) == 1 else "PEER_GATHER_FAILED", flush=True)
dist.barrier()
s.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)

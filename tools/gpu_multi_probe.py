"""Not a pytest file: N-rank probe of the fused peer flush (run under torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import rthx
from rthx.dist import ShardedTracer
from rthx._abi import RTHX_ZERO_NONE
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
name = os.environ.get("PROBE_MESH", "cfg2")
flat = rthx.flatten_domain(getattr(rthx.meshes, name)())
rpe = int(os.environ.get("PROBE_RPE", "20000"))
ref = ShardedTracer(flat, device=lr, rank=rank, world=world, n_bins=1, mode="nccl")
ref.trace(rpe, seed=5)
torch.cuda.synchronize()
ref_c = ref.counts.clone() if rank == 0 else None
N = ref.N
own = torch.arange(N, device=f"cuda:{lr}") % world
def report(tag, c):
    if rank != 0: return
    diff = (c - ref_c)
    per_owner = [int(diff[0][own == r].abs().sum().item()) for r in range(world)]
    print(tag, "total", int(c.sum().item()), "ref", int(ref_c.sum().item()), "absdiff by owner", per_owner, "min/max diff", int(diff.min().item()), int(diff.max().item()), flush=True)
sh = ShardedTracer(flat, device=lr, rank=rank, world=world, n_bins=1, mode="fused")
for rep in range(3):
    sh.trace(rpe, seed=5)
    torch.cuda.synchronize(); dist.barrier()
    report(f"A own-row memset2D rep{rep}", sh.counts if rank == 0 else None)
    dist.barrier()
# variant B: rank 0 zeroes everything, barrier, then trace without zeroing
for rep in range(2):
    if rank == 0:
        sh.counts.zero_(); sh.lost.zero_()
    torch.cuda.synchronize(); dist.barrier()
    stream = torch.cuda.current_stream(lr).cuda_stream
    sh.tracer.trace_device(rpe, sh.counts_ptr, sh.lost_ptr, stream=stream, zero_first=RTHX_ZERO_NONE, emitter_rank=rank, emitter_world=world, seed=5)
    torch.cuda.synchronize(); dist.barrier()
    report(f"B rank0-zero rep{rep}", sh.counts if rank == 0 else None)
    dist.barrier()
sh.close(); ref.close()
dist.destroy_process_group()

"""Where a ShardedTracer step spends its time (torchrun, N >= 2): CUDA events between the pieces of the fused step on every rank
(consumed-flag signal / wait, zero + trace kernel, done-flag signal, rank 0's wait for all), for the strong-scaled cfg3 step."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import rthx
from rthx.dist import ShardedTracer
from rthx._abi import RTHX_ZERO_OWN_ROWS, RTHX_DEST_PEER

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
flat = rthx.flatten_domain(rthx.meshes.cfg3())
rpe = 942951
sh = ShardedTracer(flat, device=lr, rank=rank, world=world, n_bins=1, mode="fused")
L = sh._L
st = torch.cuda.current_stream(lr)
cs = C.c_void_p(st.cuda_stream)
fl = sh._flags_ptr
W = world


def step(s, ev, use_flags=True):
    b = s & 1
    ev[0].record(st)
    if use_flags:
        if rank == 0 and s >= 1:
            L.rthx_flag_signal(C.c_void_p(fl + 8 * W), s, cs)
        if s >= 2:
            L.rthx_flag_wait(C.c_void_p(fl + 8 * W), 1, s - 1, 30.0, C.c_void_p(fl + 8 * (W + 1)), cs)
    else:
        dist.barrier(device_ids=[lr])
    ev[1].record(st)
    sh.tracer.trace_device(rpe, sh._counts_ptrs[b], sh._lost_ptrs[b], stream=st.cuda_stream, zero_first=RTHX_ZERO_OWN_ROWS | (RTHX_DEST_PEER if rank else 0), seed=50 + s,
                           emitter_rank=rank, emitter_world=world)
    ev[2].record(st)
    if use_flags:
        L.rthx_flag_signal(C.c_void_p(fl + 8 * rank), s + 1, cs)
        ev[3].record(st)
        if rank == 0:
            L.rthx_flag_wait(C.c_void_p(fl), W, s + 1, 30.0, C.c_void_p(fl + 8 * (W + 1)), cs)
    else:
        ev[3].record(st)
        dist.barrier(device_ids=[lr])
    ev[4].record(st)


for use_flags in (True, False):
    s0 = sh.step
    n = 8
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(n)]
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for k in range(n):
        step(s0 + k, evs[k], use_flags)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sh.step = s0 + n
    seg = [sum(evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(3, n)) / (n - 3) for i in range(4)]
    tot = evs[3][0].elapsed_time(evs[n - 1][4]) / (n - 3)
    print(f"rank {rank} flags={use_flags}: pre-wait {seg[0]:.3f}  zero+kernel {seg[1]:.3f}  signal {seg[2]:.3f}  post-wait {seg[3]:.3f}  | step {tot:.3f} ms, wall {1e3*(t1-t0)/n:.3f} ms/step", flush=True)
    dist.barrier()
sh.close()
dist.destroy_process_group()

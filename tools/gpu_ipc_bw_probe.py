"""Bandwidth of writes into rank 0's CUDA-IPC buffer from all other ranks at once (torchrun, N >= 2): SM stores (a torch copy
kernel on a view of the mapped memory) vs the copy engines (cudaMemcpyAsync), 112 MB per rank like one strong-scaled cfg3 step."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from rthx._lib import SharedDeviceBuffer

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
n = 14_000_000 * world        # int64 elements: 112 MB per rank
hb = torch.zeros(64, dtype=torch.uint8, device=dev)
if rank == 0:
    buf = SharedDeviceBuffer(lr, n)
    hb.copy_(torch.tensor(list(buf.handle), dtype=torch.uint8))
dist.broadcast(hb, src=0)
if rank != 0:
    buf = SharedDeviceBuffer(lr, n, handle=bytes(hb.cpu().tolist()))
view = torch.as_tensor(buf, device=dev)
per = n // world
mine = view[rank * per:(rank + 1) * per]
src = torch.ones(per, dtype=torch.int64, device=dev)
rt = C.CDLL("libcudart.so.12")
st = torch.cuda.current_stream(dev)
for name in ("sm_store", "copy_engine"):
    for rep in range(2):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for k in range(5):
            if name == "sm_store":
                mine.copy_(src)
            else:
                rt.cudaMemcpyAsync(C.c_void_p(mine.data_ptr()), C.c_void_p(src.data_ptr()), C.c_size_t(per * 8), C.c_int(4), C.c_void_p(st.cuda_stream))
        b.record(st)
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"rank {rank} {name}: {per * 8 / 1e6:.0f} MB in {ms:.3f} ms = {per * 8 / ms / 1e6:.1f} GB/s", flush=True)
    dist.barrier()
dist.barrier()
if rank != 0:
    buf.close()
dist.barrier()
if rank == 0:
    buf.close()
dist.destroy_process_group()

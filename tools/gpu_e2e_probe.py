"""Not a pytest file: where does the end-to-end time go (create / kernel / D2H / destroy)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rthx
rtm = rthx.meshes.cfg3(); flat = rthx.flatten_domain(rtm)
N = flat.n_elements
pinned = torch.empty((1, N, N), dtype=torch.int64, pin_memory=True)
torch.cuda.synchronize()
for rep in range(6):
    t0 = time.perf_counter(); tr = rthx.DeviceTracer(flat, 0); t1 = time.perf_counter()
    out = tr.trace(100000, counts_out=pinned.numpy().view(np.uint64), seed=rep); t2 = time.perf_counter()
    out2 = tr.trace(100000, counts_out=pinned.numpy().view(np.uint64), seed=rep); t3 = time.perf_counter()
    tr.close(); t5 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms | trace#1 pinned {1e3*(t2-t1):.1f} (kernel {out['stats']['kernel_ms']:.1f}, total_dev {out['stats']['total_ms']:.1f}) | "
          f"trace#2 pinned {1e3*(t3-t2):.1f} (total_dev {out2['stats']['total_ms']:.1f}) | close {1e3*(t5-t3):.1f}", flush=True)
pageable = np.empty((1, N, N), np.uint64)
tr = rthx.DeviceTracer(flat, 0)
for rep in range(3):
    t0 = time.perf_counter(); out = tr.trace(942951, counts_out=pageable, seed=rep); t1 = time.perf_counter()
    print(f"pageable full-size: wall {1e3*(t1-t0):.1f} ms, kernel {out['stats']['kernel_ms']:.1f}, total_dev {out['stats']['total_ms']:.1f}", flush=True)
ref = tr.trace(942951, counts_out=pinned.numpy().view(np.uint64), seed=2)
print("pageable == pinned:", bool(np.array_equal(pageable, ref["counts"])))

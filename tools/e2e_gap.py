"""Where does the end-to-end step lose time against the device-resident step?  (development probe)
Prints per-step wall time of Classifier.classify (host entry) beside the sum of its kernels' CUDA-event
times, and the same for the device entry."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svision_b200 import classifier as C, sites, weights

n = 10_000
clf = C.Classifier(weights.synthetic_weights(), device=0, max_batch=n)
rows = sites.make_sites_p1(n)
rd = clf.rows_to_device(rows)
rp = torch.from_numpy(rows).pin_memory()
lh = torch.empty((n,), dtype=torch.int32).pin_memory()
ph = torch.empty((n, 5), dtype=torch.float32).pin_memory()
for _ in range(5):
    clf.classify_device_calls(rd)
torch.cuda.synchronize()
K = 20
for mode in ("device", "host", "host-pageable"):
    clf.set_profiling(True); clf.profile_read(True)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(K):
        if mode == "device":
            clf.classify_device_calls(rd)
        elif mode == "host":
            clf.classify(rp.numpy(), lh.numpy(), ph.numpy())
        else:
            clf.classify(rows)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / K * 1e3
    prof = clf.profile_read(True); clf.set_profiling(False)
    ksum = sum(v[0] for v in prof.values()) / K
    print(f"{mode:14s} wall {dt:7.3f} ms/step   kernels {ksum:7.3f} ms/step   gap {dt - ksum:6.3f} ms", flush=True)
clf.close()

"""Development probe run on the GPU box: `python tools/gpu_probe.py <stage>`.
Stages are independent processes so that a device trap in one does not hide the others."""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svision_b200 import classifier as C, sites, weights  # noqa: E402


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def stage_encoder():
    from oracle import encoder as enc, encoder_c
    g = np.load(os.path.join(ROOT, "tests", "golden", "encoder_golden.npz"))
    rows, off, codes = g["rows"], g["offsets"], g["codes"]
    clf = C.Classifier(None, device=0, max_batch=256)
    ref_bits = encoder_c.encode_bits(rows)           # pinned to golden by the CPU tests
    lo = torch.tensor([l[0] for l in enc.LEVELS])
    hi = torch.tensor([l[1] for l in enc.LEVELS])
    for dt in (torch.float32, torch.float16):
        img = clf.encode(rows, dtype=dt).cpu()
        ref = torch.where(torch.from_numpy(ref_bits).permute(0, 2, 3, 1).bool(), hi, lo).to(dt)
        bad = (img != ref).reshape(img.shape[0], -1).any(1)
        print(f"encoder {dt}: rows {rows.shape[0]} mismatching images {int(bad.sum())}", flush=True)
        if bad.any():
            i = int(bad.nonzero()[0])
            d = (img[i] != ref[i]).nonzero()
            print("  first bad row", i, rows[i], "n diff", d.shape[0], d[:8].tolist())
    # timing, 16k sites
    big = sites.make_sites_p1(16384, seed=1)
    rd = clf.rows_to_device(big)
    for dt in (torch.float32, torch.float16):
        out = torch.empty((big.shape[0], 227, 227, 3), dtype=dt, device="cuda")
        ms = ev_time(lambda: clf.encode(rd, dtype=dt, out=out))
        by = big.shape[0] * (48 + 154587 * out.element_size())
        print(f"encode {dt}: {ms:.3f} ms / {big.shape[0]} sites -> {big.shape[0]/ms*1e3:.3e} sites/s, "
              f"{by/ms/1e6:.1f} GB/s algorithmic", flush=True)


def stage_cnn():
    from oracle import alexnet, encoder_c
    w = weights.synthetic_weights()
    g = np.load(os.path.join(ROOT, "tests", "golden", "cnn_golden.npz"))
    rows = g["rows"][:32]
    imgs = encoder_c.encode_f32(rows)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_logits, inter = alexnet.forward(imgs, w, torch.float32, return_intermediates=True)
    for prec in ("3pass", "1pass"):
        clf = C.Classifier(w, device=0, max_batch=32, precision=prec)
        rd = clf.rows_to_device(rows)
        labels, probs, logits = clf.classify_device(rd, want_logits=True)
        torch.cuda.synchronize()
        for name in ("norm1", "norm2", "conv3", "conv4", "pool5", "fc6", "fc7"):   # conv2 / conv5 are never materialised
            got = clf.debug_activation(name, rows.shape[0])
            ref = inter[name].numpy()
            err = np.abs(got - ref).max()
            print(f"  [{prec}] {name:6s} max abs err {err:.3e}  (ref max {np.abs(ref).max():.3f})", flush=True)
        dl = (logits.cpu().double() - torch.from_numpy(g["logits_fp64"][:32])).abs().max().item()
        ps = torch.softmax(torch.from_numpy(g["logits_fp64"][:32]), 1)
        dp = (probs.cpu().double() - ps).abs().max().item()
        lab_ok = (labels.cpu().numpy() == g["logits_fp64"][:32].argmax(1)).all()
        print(f"[{prec}] logits max err vs fp64 {dl:.3e}; softmax max err {dp:.3e}; labels equal {lab_ok}", flush=True)
        # also through svx_forward on encoder images
        im16 = clf.encode(rd, dtype=torch.float16)
        l2 = clf.forward(im16)
        print(f"[{prec}] forward(images) vs classify logits max diff {(l2 - logits).abs().max().item():.3e}", flush=True)
        got = clf.debug_activation("conv1", rows.shape[0]); ref = inter["conv1"].numpy()
        print(f"  [{prec}] dense conv1 (forward path) max abs err {np.abs(got-ref).max():.3e}", flush=True)
        clf.close()


def stage_bench():
    w = weights.synthetic_weights()
    n = int(os.environ.get("PROBE_N", 4096))
    for prec in ("3pass", "1pass"):
        clf = C.Classifier(w, device=0, max_batch=2048, precision=prec)
        rows = sites.make_sites_p1(n, seed=sites.SEED_CONFIG2)
        rd = clf.rows_to_device(rows)
        ms = ev_time(lambda: clf.classify_device(rd), iters=3, warm=2)
        print(f"[{prec}] classify_device {n} sites: {ms:.2f} ms -> {n/ms*1e3:.1f} sites/s", flush=True)
        t = time.perf_counter()
        clf.classify(rows)
        print(f"[{prec}] classify(host) {n} sites: {(time.perf_counter()-t)*1e3:.2f} ms", flush=True)
        clf.close()


def stage_counters():
    w = weights.synthetic_weights()
    n = 2048
    os.environ["SVX_DBG"] = "1"
    for prec in os.environ.get("PROBE_PREC", "3pass,1pass").split(","):
        clf = C.Classifier(w, device=0, max_batch=2048, precision=prec)
        rd = clf.rows_to_device(sites.make_sites_p1(n, seed=sites.SEED_CONFIG2))
        for _ in range(2):
            clf.classify_device(rd)
        torch.cuda.synchronize()
        clf.debug_counters(reset=True)
        iters = 3
        clf.set_profiling(True); clf.profile_read(True)
        for _ in range(iters):
            clf.classify_device(rd)
        torch.cuda.synchronize()
        prof = clf.profile_read(True)
        c = clf.debug_counters().astype(np.float64)
        names = ["conv1", "conv2", "conv3", "conv4", "conv5", "fc6", "fc7"]
        print(f"[{prec}] per-k-block cycles (avg over CTAs): total | wait operands | wait tmem | producer wait | epi: wait, drain, store (per k-block)   ms/launch")
        for i, nm in enumerate(names):
            kb = max(c[i, 3], 1)
            print(f"  {nm:6s} kb/CTA={kb/iters/148:8.0f}  total={c[i,0]/kb:7.0f}  wait_op={c[i,1]/kb:7.0f}  wait_tm={c[i,2]/kb:7.0f}  prod_wait={c[i,4]/kb:7.0f}  epi_wait={c[i,5]/kb:7.0f} drain={c[i,6]/kb:7.0f} store={c[i,7]/kb:7.0f}   {prof[nm][0]/iters:.3f} ms", flush=True)
        print("  other kernels, ms/launch: " + ", ".join(
            f"{k} {prof[k][0]/iters:.3f}" for k in ("encode", "pool2_lrn2", "pool5", "fc8_softmax")), flush=True)
        clf.close()


if __name__ == "__main__":
    {"encoder": stage_encoder, "counters": stage_counters, "cnn": stage_cnn, "bench": stage_bench}[sys.argv[1]]()

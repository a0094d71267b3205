#!/usr/bin/env python3
"""Benchmark of the SVision encode+classify hot path (BASELINE.json: candidate SV sites/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--single-process]

A *step* is one pass of the hot path over one batch of synthetic candidate sites: packed
``int32[n,12]`` rows -> 227x227x3 similarity images (never leaving the device) -> AlexNet ->
per-site (label, softmax).  Headline workload = BASELINE.json ``configs[1]``: 10 000 synthetic
sites per GPU (set P1, seed 20261017); with N>1 every rank processes its own
10 000 sites (weak scaling) and the step ends with the exchange of the per-site (label, score)
pairs, 8 B/site.  Every rank's shard carries the 256 known-answer rows of
``tests/golden/cnn_golden.npz``; outside the timed region the WHOLE gathered buffer (every rank's
slice) is compared with their golden labels / softmax.

The same run also measures ``configs[2]`` (``strong_100k``): 100 000 HiFi-profile sites (seed
20261018) strong-sharded in contiguous ceil(N/R) slices, each rank streaming its slice through the
device in balanced micro-batches with ONE exchange at the end of the stream, as the north-star
prescribes (no per-micro-batch synchronisation between ranks).

Printed JSON (one line, rank 0): the driver contract plus
  * ``value``  : sites/s with the rows already resident in HBM (CUDA events, max over ranks),
  * ``e2e``    : the same metric through the host entry ``Classifier.classify`` (C-ABI
                 ``svx_classify``): pinned host rows -> H2D -> kernels -> D2H labels+probs,
  * ``roofline``: the conv2 launch of the tensor-core layer kernel, algorithmic FLOPs / live
                 CUDA-event time, against the measured bf16 peak of MEASURED_PEAKS.json; ``traffic``
                 is parsed from the newest committed ncu summary under profiles/,
  * ``cpu_baseline``: the reference's own encoder (``BatchGenerator.next_batch`` from
                 ``baseline/_ref``, the unmodified reference installed with pip --no-deps) + a
                 torch-CPU restatement of alexnet.py (TensorFlow 1.14 is not installable here), run
                 the way the reference parallelises Step 2 (a pool of cpu_count//3 processes,
                 SVision:311-323), on a bounded sample, on this box's host cores,
  * ``per_rank``: every rank's own time per step and SM clock (names the limiter at N>1).
``--impl reference`` times that CPU path alone (the reference arm).  ``--single-process`` drives
``--gpus N`` devices from ONE process through ``svx_multi_*`` (the drop-in for the reference's single
``SVision`` process) and reports only the end-to-end number.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate SV sites/sec (encode+CNN)"
UNIT = "sites/s"
SITES_PER_GPU = 10_000
STRONG_SITES = 100_000
KNOWN = 256                                  # known-answer rows per rank shard
# sites resident on the device at once.  One micro-batch per step: measured +2 % against 2048-site
# micro-batches (fc6/fc7 get 640 tiles for 74 CTA pairs = 96 % full waves instead of 86 %, and the
# front end / finishing passes / fc8 lose their per-launch tails); needs ~15 GB of HBM
MICRO_BATCH = int(os.environ.get("SVX_BENCH_MICRO_BATCH", 10_000))
WORKLOAD = ("configs[1]: synthetic 10k candidate sites per GPU (set P1, seed 20261017), "
            "227x227x3 images, encode+CNN")
CNN_FLOP_PER_SITE = 1_440_662_592            # SURVEY.md 8(a) layer table (2 x 720 331 296 MACs)
ENC_BYTES_PER_SITE = 48 + 227 * 227 * 3 * 2  # 16-bit image is what is emitted
#: algorithmic FLOPs per site of each tensor-core layer (groups honoured, no padding credit)
LAYER_FLOP = {"conv1": 2 * 105_415_200, "conv2": 2 * 223_948_800, "conv3": 2 * 149_520_384,
              "conv4": 2 * 112_140_288, "conv5": 2 * 74_760_192, "fc6": 2 * 37_748_736,
              "fc7": 2 * 16_777_216}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


def conv2_traffic_from_profiles():
    """(bytes per site, source file) of the conv2 launch from the newest committed `ncu --set full`
    summary under profiles/ (tools/ncu_summary.py format): dram__bytes_read.sum + dram__bytes_write.sum
    of the first 128-column layer-kernel launch, per site of that capture.  None if absent."""
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_full_summary*.txt"))):
        sites = None
        for line in open(path):
            m = re.search(r"sites per launch[:=]\s*(\d+)", line)
            if m:
                sites = int(m.group(1))
            if re.match(r"\s*(layer_tc_kernel|conv_tc2_kernel)<128, 3", line) and sites:
                f = line.split()
                try:
                    i = next(k for k, tok in enumerate(f) if tok.startswith("("))      # grid column
                    rd, wr = float(f[i + 7]), float(f[i + 8])                          # dramR[MB], dramW[MB]
                except (StopIteration, ValueError, IndexError):
                    continue
                best = ((rd + wr) * 1e6 / sites, os.path.basename(path))
                break
    return best


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (profiling recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            sm, mx, reasons = [], [], set()
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
            if sm:
                out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                       "reasons": sorted(reasons), "samples": len(sm)}
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def load_known_answers():
    g = np.load(os.path.join(ROOT, "tests", "golden", "cnn_golden.npz"))
    logits = g["logits_fp64"][:KNOWN]
    e = np.exp(logits - logits.max(1, keepdims=True))
    probs = e / e.sum(1, keepdims=True)
    labels = logits.argmax(1).astype(np.int32)
    return g["rows"][:KNOWN].astype(np.int32), labels, probs[np.arange(KNOWN), labels]


def known_slots(per: int) -> np.ndarray:
    """Where the known-answer rows sit inside a rank's shard of `per` rows: half at the head, half in
    the last slots."""
    half = KNOWN // 2
    return np.concatenate([np.arange(half), np.arange(per - half, per)])


def check_known(labels: np.ndarray, scores: np.ndarray, world: int, per: int, k_labels, k_scores):
    """Every rank's slice of the gathered (labels, scores) against the golden answers."""
    slots = known_slots(per)
    ok, worst = True, 0.0
    for r in range(world):
        idx = r * per + slots
        ok &= bool(np.array_equal(labels[idx], k_labels))
        worst = max(worst, float(np.abs(scores[idx].astype(np.float64) - k_scores).max()))
    return bool(ok and worst < 1e-3), worst


# ------------------------------------------------------------------------------------------------
# CPU baseline = the reference's own path on this box's host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def _ref_worker_init(bed_path, weights_npz, torch_threads, use_ref):
    """Pool initializer: one reference-style worker = its own generator over its own BED + its own
    copy of the model (Predict.run builds both per process: predict.py:155-189).  Not timed."""
    import torch
    torch.set_num_threads(torch_threads)
    from oracle import alexnet
    w = dict(np.load(weights_npz))
    _W["alexnet"], _W["weights"], _W["torch"] = alexnet, w, torch
    _W["bed"] = bed_path
    _W["use_ref"] = use_ref
    if use_ref:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        from src.network.create_batch import BatchGenerator      # the UNMODIFIED reference
        _W["BatchGenerator"] = BatchGenerator
    else:
        from oracle import encoder_c
        encoder_c.build()
        encoder_c.set_threads(torch_threads)
        _W["encoder_c"] = encoder_c


def _ref_worker_run(task):
    """Encode + classify rows [lo, hi) of the worker's BED, batch by batch as predict.py:206-210."""
    idx, lo, hi, batch = task
    torch, alexnet, w = _W["torch"], _W["alexnet"], _W["weights"]
    n = hi - lo
    if n <= 0:
        return 0
    if _W["use_ref"]:
        path = f"{_W['bed']}.{idx}.bed"
        # as predict.py:162-164 builds it; parses the slice's BED and pads it to whole batches
        gen = _W["BatchGenerator"](path, horizontal_flip=False, shuffle=False, nb_classes=5, batch_size=batch)
        done = 0
        while done < n:
            images, _ = gen.next_batch(batch)                     # create_batch.py:88-155
            logits = alexnet.forward(np.ascontiguousarray(images, dtype=np.float32), w, torch.float32)
            torch.softmax(logits, 1), torch.argmax(logits, 1)
            done += batch
    else:
        rows = np.load(f"{_W['bed']}.rows.npy")[lo:hi]
        for s in range(0, n, batch):
            images = _W["encoder_c"].encode_f32(rows[s:s + batch])
            logits = alexnet.forward(images, w, torch.float32)
            torch.softmax(logits, 1), torch.argmax(logits, 1)
    return n


class CpuReference:
    """The reference's CPU path, parallelised the way the reference does it: a pool of
    max(1, cores // 3) processes (SVision:311-323), each encoding and classifying its own slice in
    batches of 128 (SVision:88).  Encoder: the real ``BatchGenerator`` when ``baseline/_ref`` is
    importable (kind "reference+proxy"), else the oracle's C port (kind "port").  CNN: torch-CPU
    restatement of alexnet.py, 3 threads per process (TensorFlow 1.14 is not installable here)."""

    def __init__(self, weights):
        import multiprocessing as mp
        self.cores = os.cpu_count() or 1
        self.procs = max(1, self.cores // 3)
        self.threads = max(1, self.cores // self.procs)
        self.use_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "src", "network", "create_batch.py"))
        try:
            import cv2  # noqa: F401  (the reference's rasteriser)
        except Exception:
            self.use_ref = False
        self.kind = "reference+proxy" if self.use_ref else "port"
        self.tmp = tempfile.mkdtemp(prefix="svx_ref_")
        self.base = os.path.join(self.tmp, "sample")
        np.savez(os.path.join(self.tmp, "w.npz"), **weights)
        self.pool = mp.get_context("spawn").Pool(
            self.procs, initializer=_ref_worker_init,
            initargs=(self.base, os.path.join(self.tmp, "w.npz"), self.threads, self.use_ref))

    def describe(self) -> str:
        enc = ("the reference's own BatchGenerator.next_batch (unmodified install under baseline/_ref)"
               if self.use_ref else "C restatement of the encoder (oracle/encoder_c.c)")
        return (f"{enc} + torch-CPU fp32 restatement of alexnet.py (proxy for TensorFlow 1.14 CPU, not "
                f"installable here); {self.procs} processes x {self.threads} threads as SVision:311-323, batch 128")

    def run(self, rows: np.ndarray, batch: int = 128):
        """(sites/s, seconds) for `rows`; writing the slice BEDs is not timed (text I/O
        is outside the metric), parsing them inside BatchGenerator is."""
        from svision_b200 import sites
        n = rows.shape[0]
        per = -(-n // self.procs)
        per = -(-per // batch) * batch                           # whole batches per worker
        tasks = []
        if not self.use_ref:
            np.save(f"{self.base}.rows.npy", rows)
        for i in range(self.procs):
            lo, hi = min(n, i * per), min(n, (i + 1) * per)
            if self.use_ref and hi > lo:
                with open(f"{self.base}.{i}.bed", "w") as f:
                    f.write("\n".join(sites.rows_to_bed_lines(rows[lo:hi])) + "\n")
            tasks.append((i, lo, hi, batch))
        t0 = time.perf_counter()
        done = sum(self.pool.map(_ref_worker_run, tasks, chunksize=1))
        dt = time.perf_counter() - t0
        assert done == n
        return n / dt, dt

    def close(self):
        try:
            self.pool.close()
            self.pool.join()
        finally:
            import shutil
            shutil.rmtree(self.tmp, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from svision_b200 import sites, weights
    w = weights.synthetic_weights()
    rows_all = sites.make_sites_p1(SITES_PER_GPU, seed=sites.SEED_CONFIG2)
    ref = CpuReference(w)
    # step size: the whole 10 000-site workload when the run then still ends within ~5 minutes,
    # otherwise a bounded sample of it (the rate does not depend on the sample: sites are independent
    # and every image costs the same dense CNN)
    ref.run(rows_all[:1024])                             # cold: worker start-up, first-touch of the weights
    rate0, _ = ref.run(rows_all[:2048])
    steps_total = args.steps + args.warmup
    sample = int(os.environ.get("SVX_REF_SAMPLE", 0)) or int(min(SITES_PER_GPU, max(1024, rate0 * 300 / steps_total)))
    sample = min(SITES_PER_GPU, -(-sample // 128) * 128) if sample < SITES_PER_GPU else SITES_PER_GPU
    for i in range(args.warmup):
        ref.run(rows_all[:sample])
    t0 = time.perf_counter()
    for i in range(args.steps):
        o = 0 if sample == SITES_PER_GPU else (i * sample) % (SITES_PER_GPU - sample)
        ref.run(rows_all[o:o + sample])
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    step_note = ("the whole workload per step" if sample == SITES_PER_GPU
                 else f"bounded sample of {sample} sites of that workload per step (rate is sample-independent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "sites_per_step": sample, "step": step_note},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                         "sample": f"{sample} sites/step x {args.steps} steps; " + ref.describe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    ref.close()
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
def run_single_process(args):
    """All `--gpus` devices driven from ONE process through svx_multi_* (the reference's single
    SVision process, SVision:296-341): host rows in, host labels+probs out, dynamic chunk balancing."""
    import torch
    from svision_b200 import classifier as C, sites, weights
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    ndev = min(args.gpus, torch.cuda.device_count())
    k_rows, k_labels, k_scores = load_known_answers()
    n = STRONG_SITES
    rows = sites.make_sites_p1(n, seed=sites.SEED_CONFIG3).copy()
    slots = known_slots(n)
    rows[slots] = k_rows
    rows_pinned = torch.from_numpy(rows).pin_memory()
    labels = torch.empty((n,), dtype=torch.int32).pin_memory()
    probs = torch.empty((n, 5), dtype=torch.float32).pin_memory()
    multi = C.MultiClassifier(weights.synthetic_weights(), devices=list(range(ndev)), max_batch=MICRO_BATCH)
    for _ in range(max(args.warmup, 3)):
        multi.classify(rows_pinned.numpy(), labels.numpy(), probs.numpy())
    t0 = time.perf_counter()
    for _ in range(args.steps):
        multi.classify(rows_pinned.numpy(), labels.numpy(), probs.numpy())
    dt = time.perf_counter() - t0
    l, p = labels.numpy(), probs.numpy()
    ok, worst = check_known(l, p[np.arange(n), l], 1, n, k_labels, k_scores)
    value = n * args.steps / dt
    emit({"impl": "b200-single-process", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ndev,
          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3,
          "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
          "dtype": "f16x3 (fp16 hi/lo split operands, fp32 accumulate)", "data": "synthetic",
          "config": {"workload": "configs[2]: synthetic 100k candidate sites, HiFi profile (set P1, seed 20261018), "
                                 "one process driving all devices through svx_multi_classify",
                     "micro_batch": MICRO_BATCH, "last_split_sites_per_device": multi.last_split()},
          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": n * 48, "d2h_bytes_per_step": n * 24},
          "parity_spot_check": {"ok": ok, "known_answers_checked": KNOWN, "max_abs_score_err": worst}})
    multi.close()
    return 0


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from svision_b200 import classifier as C, sharded, sites, weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = weights.synthetic_weights()
    clf = C.Classifier(w, device=local, max_batch=MICRO_BATCH, precision=args.precision)
    k_rows, k_labels, k_scores = load_known_answers()
    n = SITES_PER_GPU
    rows_all = sites.make_sites_p1(n * world, seed=sites.SEED_CONFIG2)
    rows = np.ascontiguousarray(rows_all[rank * n:(rank + 1) * n]).copy()      # contiguous shard
    rows[known_slots(n)] = k_rows                                               # known answers in EVERY shard
    rows_dev = clf.rows_to_device(rows)
    rows_pinned = torch.from_numpy(rows).pin_memory()
    labels_host = torch.empty((n,), dtype=torch.int32).pin_memory()
    probs_host = torch.empty((n, 5), dtype=torch.float32).pin_memory()
    gathered = torch.empty((world * n, 2), dtype=torch.int32, device=dev) if world > 1 else None
    # how the per-site (label, score) calls -- what predict.py:230,251 consumes -- reach every rank:
    #   fused: the fc8 kernel stores them into every rank's buffer over NVLink (svx_classify_exchange;
    #          default -- measured equal to NCCL within noise, no collective kernel on the path)
    #   nccl : ONE all-gather of 8 B/site (the north-star's prescription; SVX_BENCH_EXCHANGE=nccl)
    exchange_mode = os.environ.get("SVX_BENCH_EXCHANGE", "fused") if world > 1 else "none"
    exchange, exchange_note = None, ""
    if exchange_mode == "fused":
        try:
            exchange = sharded.Exchange(clf, n)          # raises on EVERY rank if any rank fails
        except Exception as ex:                          # noqa: BLE001 -- e.g. no peer access between the GPUs
            exchange_mode = "nccl"
            exchange_note = f" (fused exchange unavailable: {type(ex).__name__}: {ex})"[:200]

    def step_device():
        """One step: this rank's rows -> per-site calls (svx_call = int32 label, fp32 score, written
        by the fc8 kernel), gathered on every rank.  No torch compute kernel runs in the step."""
        if exchange is not None:
            return exchange.classify(rows_dev)
        calls = clf.classify_device_calls(rows_dev, raw=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, calls)
            calls = gathered
        return calls[:, 0], calls[:, 1].view(torch.float32)

    rows_stage = torch.empty_like(rows_dev)
    calls_host = torch.empty((world * n, 2), dtype=torch.int32).pin_memory() if world > 1 else None

    def step_e2e():
        """End to end from pinned host rows to host results.  One GPU: the host entry svx_classify (labels +
        probs back).  Several: this step's rows H2D, classify + exchange, and the gathered calls of ALL ranks
        D2H on every rank (what each rank's downstream replay would read)."""
        if world == 1:
            clf.classify(rows_pinned.numpy(), labels_host.numpy(), probs_host.numpy())
            return
        rows_stage.copy_(rows_pinned, non_blocking=True)
        if exchange is not None:
            calls = exchange.classify(rows_stage, raw=True)
        else:
            dist.all_gather_into_tensor(gathered, clf.classify_device_calls(rows_stage, raw=True))
            calls = gathered
        calls_host.copy_(calls, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    # ---- warm-up (>= 3) --------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    step_e2e()
    barrier()

    # ---- timed: device-resident inputs --------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()                                      # every rank samples its own GPU
    clf.set_profiling(True)
    clf.profile_read(reset=True)
    C.launch_count(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    my_ms = e0.elapsed_time(e1)
    ms_total = max_over_ranks(my_ms)
    launches = C.launch_count()
    prof = clf.profile_read(reset=True)
    clf.set_profiling(False)
    # this rank's own kernel time per step (without waiting for the other ranks' flags)
    my_kernel_ms = sum(v[0] for v in prof.values()) / args.steps

    # ---- timed: end to end through the host entry ---------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()
    per_rank = all_ranks({"rank": rank, "ms_per_step": my_ms / args.steps, "kernel_ms_per_step": my_kernel_ms,
                          "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")})

    # ---- parity of what was timed (not timed itself): the WHOLE gathered buffer, every rank's slice,
    #      against the known answers; plus device entry == host entry, bit for bit, on this rank's shard
    l_step, s_step = step_device()
    torch.cuda.synchronize()
    if exchange is not None:
        exchange.status()
    l_all, s_all = l_step.cpu().numpy(), s_step.cpu().numpy()
    known_ok, known_err = check_known(l_all, s_all, world, n, k_labels, k_scores)
    if world == 1:
        lh, ph = labels_host.numpy(), probs_host.numpy()
        bits_ok = bool(np.array_equal(l_all, lh)) and bool(np.array_equal(s_all, ph[np.arange(n), lh]))
    else:                                                # what the e2e steps brought to the host == this step's
        ch = calls_host.numpy()
        bits_ok = bool(np.array_equal(l_all, ch[:, 0])) and bool(np.array_equal(s_all.view(np.int32), ch[:, 1]))
    parity_all = all_ranks({"known": known_ok, "err": known_err, "bits": bits_ok})

    # ---- configs[2]: 100 k HiFi sites strong-sharded, ONE exchange at the end of the stream --------
    strong = run_strong(args, clf, sharded, sites, dist, world, rank, dev, k_rows, k_labels, k_scores,
                        barrier, max_over_ranks, all_ranks, exchange is not None)

    # ---- standalone encoder (svx_encode, 16-bit NHWC images to HBM): its HBM roofline, not timed above
    enc = None
    if rank == 0:
        n_enc = min(n, 8192)
        img = torch.empty((n_enc, 227, 227, 3), dtype=torch.float16, device=dev)
        for _ in range(3):
            clf.encode(rows_dev[:n_enc], dtype=torch.float16, out=img)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        a.record()
        for _ in range(iters):
            clf.encode(rows_dev[:n_enc], dtype=torch.float16, out=img)
        b.record()
        torch.cuda.synchronize()
        enc = (a.elapsed_time(b) / iters, n_enc)
        del img

    if rank == 0:
        peaks = load_peaks()
        total_sites = n * world * args.steps
        value = total_sites / (ms_total * 1e-3)
        gemm_slots = tuple(k for k in ("conv1", "conv2", "conv3", "conv4", "conv5", "fc6", "fc7")
                           if prof[k][1] > 0)
        gemm_ms = sum(prof[k][0] for k in gemm_slots)
        gemm_launches = sum(prof[k][1] for k in gemm_slots)
        sites_rank = n * args.steps
        flops = sum(LAYER_FLOP[k] for k in gemm_slots) * sites_rank
        achieved = flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        enc_ms = prof["encode"][0]
        enc_gbs = ENC_BYTES_PER_SITE * sites_rank / (enc_ms * 1e-3) / 1e9 if enc_ms > 0 else 0.0
        layers = {}
        for k in C.Classifier.PROFILE_SLOTS:
            ms, cnt = prof[k]
            entry = {"ms_per_launch": ms / cnt if cnt else None, "launches": cnt,
                     "share": ms / (ms_total) if ms_total > 0 else None}
            if k in LAYER_FLOP and ms > 0:
                entry["tflops_algorithmic"] = LAYER_FLOP[k] * sites_rank / (ms * 1e-3) / 1e12
            layers[k] = entry
        # CPU baseline on a bounded sample (rank 0, N=1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = CpuReference(w)
            # ~10-20 s of CPU work on a 16-core host (the contract asks for 10-30 s)
            sample = min(int(os.environ.get("SVX_CPU_SAMPLE", 8192)), rows.shape[0])
            ref.run(rows[:512])
            rate, secs = ref.run(rows[:sample])
            cpu = {"value": rate, "unit": UNIT, "cores": ref.cores, "kind": ref.kind,
                   "sample": f"first {sample} sites of the workload, {secs:.1f} s; " + ref.describe()}
            ref.close()
        traffic = conv2_traffic_from_profiles()
        conv2_launch_sites = sites_rank / max(prof["conv2"][1], 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16x3 (fp16 hi/lo split operands, fp32 accumulate)" if args.precision == "3pass"
                     else "f16 (single pass; NOT parity-clean)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sites_per_gpu": n, "micro_batch": MICRO_BATCH, "precision": args.precision,
                       "known_answers": f"{KNOWN} rows of tests/golden/cnn_golden.npz replace the first/last "
                                        f"{KNOWN // 2} rows of every rank's shard (parity_spot_check)",
                       "weights": "synthetic He-init, calibrated fc8 (seed 1234)",
                       "collective": {"none": "none", "nccl": "one NCCL all_gather of svx_call (label, score), 8 B/site",
                                      "fused": "fused: fc8 kernel stores svx_call (8 B/site) into every rank's "
                                               "buffer over NVLink + flag barrier (svx_classify_exchange)"}[exchange_mode] + exchange_note,
                       "l2": f"activation working set per micro-batch ~{1.4e-3 * min(MICRO_BATCH, n):.1f} GB and "
                             "fp16 hi/lo weights 226 MB both exceed the 126 MB L2; no flush needed"},
            "e2e": {"value": total_sites / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(n * world * 48),
                    "d2h_bytes_per_step": int(n * 24) if world == 1 else int(world * world * n * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # dominant kernel = the conv2 launch of the tensor-core layer kernel (largest single launch)
            "roofline": {"kernel": "layer_tc_kernel<128,3,pooled> (tcgen05 cta_group::2 layer kernel), conv2 launch "
                                   "(+ ReLU + pool2 in its epilogue)",
                         "bound": "tensor",
                         "achieved": layers["conv2"].get("tflops_algorithmic", 0.0), "peak": peak,
                         "unit": "TFLOP/s",
                         "frac": layers["conv2"].get("tflops_algorithmic", 0.0) / peak if peak else None,
                         "traffic": traffic[0] * conv2_launch_sites if traffic else None,
                         "traffic_note": (f"bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum of the conv2 "
                                          f"launch in profiles/{traffic[1]} (ncu --set full), per site, times this "
                                          f"run's sites per launch") if traffic else
                                         "no ncu --set full summary of this kernel under profiles/",
                         "ms_per_launch": layers["conv2"]["ms_per_launch"],
                         "share_of_step": prof["conv2"][0] / ms_total if ms_total > 0 else None,
                         "peak_source": peaks["source"] + ", bf16 sustained",
                         "all_tensor_layers": {"layers": list(gemm_slots), "achieved": achieved,
                                               "frac": achieved / peak if peak else None,
                                               "launches": gemm_launches,
                                               "share_of_step": gemm_ms / ms_total if ms_total > 0 else None},
                         "algorithmic_gflop_per_site": sum(LAYER_FLOP[k] for k in gemm_slots) / 1e9,
                         "note": "the 3-pass fp16-split parity recipe executes 3x the algorithmic FLOPs and the "
                                 "padded grids another 1.15x (29^2/27^2, 14^2/13^2): frac counts algorithmic "
                                 "FLOPs only, so 1/3.46 = 0.29 of a same-clock peak is the ceiling by "
                                 "construction, 0.33 for the layers with less padding (ncu: tensor pipe 94-99.8 % "
                                 "active, profiles/README.md); conv1 (0.211 GFLOP/site) runs in the fused front end"},
            "encoder_roofline": None if enc is None else {
                "kernel": "encode_kernel<fp16 NHWC> (svx_encode: rows -> 227x227x3 16-bit images in HBM)",
                "bound": "hbm", "achieved": ENC_BYTES_PER_SITE * enc[1] / (enc[0] * 1e-3) / 1e9,
                "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ENC_BYTES_PER_SITE * enc[1] / (enc[0] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "ms_per_launch": enc[0], "sites_per_launch": enc[1],
                "note": "algorithmic bytes = 48 B row in + 227*227*3*2 B image out per site; measured "
                        "outside the timed step (the classify path never materialises the image)"},
            "front_end": {"kernel": "front_kernel (encode + conv1 + ReLU + pool1 + LRN1 fused; the "
                                    "image never reaches HBM)",
                          "ms_per_launch": prof["encode"][0] / max(prof["encode"][1], 1),
                          "equivalent_image_GBps": enc_gbs},
            "layers": layers,
            "per_rank": per_rank,
            "strong_100k": strong,
            "cpu_baseline": cpu,
            "parity_spot_check": {
                "ok": all(p["known"] and p["bits"] for p in parity_all),
                "known_answers_checked": world * KNOWN,
                "what": "every rank's slice of the gathered (label, score) buffer against the golden labels / "
                        "softmax of tests/golden/cnn_golden.npz (checked on every rank); what the end-to-end "
                        "steps brought to the host == the device-resident step, bit for bit",
                "max_abs_score_err": max(p["err"] for p in parity_all),
                "host_vs_device_bits": all(p["bits"] for p in parity_all)},
        }
        emit(line)
    if exchange is not None:
        exchange.close()
    clf.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_strong(args, clf, sharded, sites, dist, world, rank, dev, k_rows, k_labels, k_scores,
               barrier, max_over_ranks, all_ranks, fused):
    """BASELINE.json configs[2] / SURVEY 8(d) config 3: 100 000 P1-HiFi sites (seed 20261018) in
    contiguous shards of ceil(N/R); every rank streams its shard through the device (balanced
    micro-batches inside ONE library call) and the per-site calls are exchanged ONCE, at the end of the
    stream: no rank waits for another inside the stream."""
    import torch
    n_total = STRONG_SITES
    rows_all = sites.make_sites_p1(n_total, seed=sites.SEED_CONFIG3).copy()
    start, stop, per = sharded.shard_bounds(n_total, world, rank)
    mine = sharded.shard_rows(rows_all, world, rank)
    mine[known_slots(per)] = k_rows                      # known answers in every shard (pad rows of the last one too)
    mine_dev = clf.rows_to_device(mine)
    gathered = torch.empty((world * per, 2), dtype=torch.int32, device=dev) if world > 1 else None
    x = None
    if fused and world > 1:
        try:
            x = sharded.Exchange(clf, per)
        except Exception:                                # noqa: BLE001
            x = None

    def step():
        if x is not None:
            return x.classify(mine_dev)
        calls = clf.classify_device_calls(mine_dev, raw=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered, calls)
            calls = gathered
        return calls[:, 0], calls[:, 1].view(torch.float32)

    steps = max(2, min(args.steps, 5))
    step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    my_ms = e0.elapsed_time(e1) / steps
    ms = max_over_ranks(my_ms)
    l, s = step()
    torch.cuda.synchronize()
    if x is not None:
        x.status()
    l_host, s_host = l.cpu().numpy(), s.cpu().numpy()    # views of the exchange's buffer: copy before closing it
    if x is not None:
        x.close()
    ok, err = check_known(l_host, s_host, world, per, k_labels, k_scores)
    per_rank = all_ranks({"rank": rank, "ms_per_step": my_ms})
    oks = all_ranks(ok)
    return {"workload": "configs[2]: synthetic 100k candidate sites, HiFi profile (set P1, seed 20261018), "
                        "contiguous shards of ceil(N/R), one exchange at the end of the stream",
            "value": n_total / (ms * 1e-3), "unit": UNIT, "scaling": "strong", "sites": n_total,
            "sites_per_rank": per, "steps": steps, "ms_per_step": ms,
            "exchange": "fused (svx_classify_exchange, flag published by the last micro-batch's fc8 kernel)"
                        if x is not None else ("one NCCL all_gather of 8 B/site" if world > 1 else "none"),
            "per_rank_ms_per_step": [p["ms_per_step"] for p in per_rank],
            "parity_ok": all(oks), "known_answers_checked": world * KNOWN, "max_abs_score_err": err}


_JSON_OUT = sys.stdout


def emit(line: dict) -> None:
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def main():
    # stdout carries exactly ONE line, the JSON: anything a library prints to fd 1 (NCCL's version
    # banner on the GPU box, warnings) is sent to stderr instead
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="3pass", choices=["3pass", "1pass"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-process", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.single_process:
        return run_single_process(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())

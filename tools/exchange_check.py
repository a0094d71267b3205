"""Multi-GPU check of the fused result exchange, run under torchrun on one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/exchange_check.py [--sites 10000] [--iters 20]

Every rank classifies its contiguous shard of one synthetic stream three ways and compares them on
the device, bit for bit:  (a) svx_classify_device_calls + ONE NCCL all-gather of the 8-byte calls
(the north-star's prescription), (b) svx_classify_exchange (fc8 kernel stores the calls into every
rank's buffer over NVLink, flag barrier; no collective kernel), (c) rank 0's single-GPU result for
the whole stream.  Then both exchanges are timed (CUDA events, max over ranks).  Prints one JSON
line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svision_b200 import classifier as C, sharded, sites, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=10_000, help="sites per rank")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--timeout-test", action="store_true",
                    help="the last rank shows up late once: the others must see the timeout (sticky error, "
                         "poisoned calls), not stale results")
    a = ap.parse_args()
    if a.timeout_test:
        os.environ["SVX_EXCHANGE_TIMEOUT_MS"] = "400"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

    clf = C.Classifier(weights.synthetic_weights(), device=local, max_batch=2048)
    n_total = a.sites * world - 37                        # ragged: the last shard is padded
    rows = sites.make_sites_p1(n_total, seed=sites.SEED_CONFIG3)
    mine = clf.rows_to_device(sharded.shard_rows(rows, world, rank))
    per = mine.shape[0]
    x = sharded.Exchange(clf, per)
    gathered = torch.empty((world * per, 2), dtype=torch.int32, device=dev)

    def via_nccl():
        dist.all_gather_into_tensor(gathered, clf.classify_device_calls(mine, raw=True))
        return gathered[:, 0], gathered[:, 1].view(torch.float32)

    ok = True
    for _ in range(3):                                    # both buffer parities
        l_n, s_n = via_nccl()
        l_f, s_f = x.classify(mine)
        ok &= bool(torch.equal(l_n, l_f)) and bool(torch.equal(s_n, s_f))
        torch.cuda.synchronize()
        x.status()
    if rank == 0:                                         # the whole stream on one GPU
        l_1, s_1 = clf.classify_device_calls(clf.rows_to_device(rows))
        ok &= bool(torch.equal(l_1, l_f[:n_total])) and bool(torch.equal(s_1, s_f[:n_total]))

    timeout_ok = None
    if a.timeout_test:
        import time
        from svision_b200 import _lib
        dist.barrier()
        torch.cuda.synchronize()
        late = world - 1
        if rank == late:
            time.sleep(2.5)                               # > SVX_EXCHANGE_TIMEOUT_MS
        l_t, s_t = x.classify(mine)
        torch.cuda.synchronize()
        if rank == late:
            timeout_ok = True                             # its own wait saw everybody (they came first)
            try:
                x.status()
            except _lib.SvxError:
                timeout_ok = False
        else:
            seen, sticky = False, False
            try:
                x.classify(mine)                          # the error is sticky: refused before any launch
            except _lib.SvxError:
                sticky = True
            try:
                x.status()
            except _lib.SvxError as e:
                seen = f"rank {late}" in str(e)
            region = l_t[late * per:(late + 1) * per]
            poisoned = bool((region == -1).all().item())
            own = bool(torch.equal(l_t[rank * per:(rank + 1) * per], l_n[rank * per:(rank + 1) * per]))
            timeout_ok = seen and sticky and poisoned and own
        dist.barrier()
        # after status() the exchange works again: the refused call did not consume an epoch, so the
        # ranks are still aligned
        l_n, s_n = via_nccl()
        l_f, s_f = x.result(mine)
        timeout_ok = timeout_ok and bool(torch.equal(l_n, l_f)) and bool(torch.equal(s_n, s_f))
        ok &= timeout_ok

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / a.iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_nccl = timed(via_nccl)
    ms_fused = timed(lambda: x.classify(mine))
    x.status()
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "sites_per_rank": per, "all_paths_bit_identical": bool(flag.item()),
                          "timeout_test": "detected, poisoned, recovered on every rank" if a.timeout_test and flag.item()
                          else ("FAILED" if a.timeout_test else "not run"),
                          "ms_per_step_nccl_allgather": ms_nccl, "ms_per_step_fused_exchange": ms_fused,
                          "sites_per_s_nccl": world * per / ms_nccl * 1e3,
                          "sites_per_s_fused": world * per / ms_fused * 1e3}), flush=True)
    x.close()
    clf.close()
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())

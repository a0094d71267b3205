"""Multi-GPU sharding of the path: sites are independent, so rank r of R gets
the contiguous slice ``[r*ceil(N/R), min(N, (r+1)*ceil(N/R)))``, padded with the reference's pad
row (``create_batch.py:55``) to equal length, and the ranks exchange ONE all-gather of per-site
results at the end.  Contiguity preserves file order, which the region-flush logic of
``src/network/predict.py:235-247`` depends on.  One process per GPU (``torch.distributed``; NCCL
on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .sites import PAD_ROW


def shard_bounds(n: int, world: int, rank: int):
    """(start, stop, per_rank) of rank's contiguous shard."""
    per = -(-n // world) if n else 0
    start = min(n, rank * per)
    return start, min(n, start + per), per


def shard_rows(rows: np.ndarray, world: int, rank: int) -> np.ndarray:
    """This rank's shard, padded with PAD_ROW to ``ceil(N/R)`` rows (all_gather needs equal counts)."""
    start, stop, per = shard_bounds(rows.shape[0], world, rank)
    out = np.tile(PAD_ROW, (per, 1)).astype(np.int32)
    out[:stop - start] = rows[start:stop]
    return out


def gather_results(labels: torch.Tensor, probs: torch.Tensor, n_total: int, group=None):
    """All-gather per-rank ``labels int32[per]`` / ``probs float32[per,5]`` into the full
    ``[n_total]`` / ``[n_total,5]`` on every rank.  One collective: labels travel bit-cast inside
    the float32 payload (6 x 4 B per site; the (label, score) pair the reference consumes is 8 B)."""
    world = dist.get_world_size(group)
    per = labels.shape[0]
    payload = torch.cat([labels.view(torch.float32).reshape(per, 1), probs], dim=1).contiguous()
    out = torch.empty((world * per, payload.shape[1]), dtype=torch.float32, device=payload.device)
    dist.all_gather_into_tensor(out, payload, group=group)
    full_labels = out[:, 0].contiguous().view(torch.int32)[:n_total]
    full_probs = out[:, 1:][:n_total].contiguous()
    return full_labels, full_probs


def classify_sharded(classify_fn, rows: np.ndarray, device=None, group=None):
    """``classify_fn(rows int32[m,12]) -> (labels int32[m], probs float32[m,5])`` (numpy) is run on
    this rank's shard; returns the full-length numpy results on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = rows.shape[0]
    mine = shard_rows(rows, world, rank)
    labels, probs = classify_fn(mine)
    dev = device if device is not None else torch.device("cpu")
    l = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).to(dev)
    p = torch.from_numpy(np.ascontiguousarray(probs, dtype=np.float32)).to(dev)
    fl, fp = gather_results(l, p, n, group)
    return fl.cpu().numpy(), fp.cpu().numpy()

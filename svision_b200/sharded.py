"""Multi-GPU sharding of the path: sites are independent, so rank r of R gets
the contiguous slice ``[r*ceil(N/R), min(N, (r+1)*ceil(N/R)))``, padded with the reference's pad
row (``create_batch.py:55``) to equal length, and the ranks exchange ONE all-gather of per-site
results at the end.  Contiguity preserves file order, which the region-flush logic of
``src/network/predict.py:235-247`` depends on.  One process per GPU (``torch.distributed``; NCCL
on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import ctypes

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


class ShardedClassifier:
    """``classify(rows)`` over all ranks: every rank holds its own classifier (one GPU each), takes the
    contiguous shard of ``rows`` that :func:`shard_rows` gives it and gets the full-length results back
    (:func:`classify_sharded`: one all-gather).  Every rank must make the same calls with the same rows
    -- ``svision_b200.step2`` under ``torchrun`` does."""

    def __init__(self, local, group=None):
        self.local, self.group = local, group
        dev = getattr(local, "torch_device", None)
        self.device = dev if (dev is not None and dist.get_backend(group) == "nccl") else torch.device("cpu")

    def classify(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        if rows.shape[0] == 0:
            return np.empty((0,), np.int32), np.empty((0, 5), np.float32)
        return classify_sharded(self.local.classify, rows, device=self.device, group=self.group)


class _DeviceView:
    """A library-owned device buffer exposed through ``__cuda_array_interface__`` (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


class Exchange:
    """Fused result exchange (``include/svx.h``, SURVEY.md §8(e) "fused variant"): the fc8 kernel of
    every rank stores its per-site ``(label, score)`` calls straight into the gathered buffer of
    EVERY rank over NVLink and publishes a flag; no collective kernel runs on the path.

    ``torch.distributed`` is used once, at construction, to all-gather the CUDA IPC handle blobs.
    Every rank must call :meth:`classify` the same number of times, with at most ``sites_per_rank``
    rows (use :func:`shard_rows`)."""

    def __init__(self, classifier, sites_per_rank: int, group=None):
        from . import _lib
        self._lib = _lib.load()
        self._check = _lib.check
        self.clf = classifier
        self.per = int(sites_per_rank)
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if distributed else 1
        self.rank = dist.get_rank(group) if distributed else 0
        self._x = None
        err = None
        x = ctypes.c_void_p()
        mine = np.zeros(_lib.IPC_HANDLE_BYTES + 1, dtype=np.uint8)      # handle + "this rank is fine"
        try:
            self._check(self._lib.svx_exchange_create(classifier._h, self.rank, self.world, self.per,
                                                      ctypes.byref(x)), "svx_exchange_create")
            self._x = x
            if self.world > 1:
                self._check(self._lib.svx_exchange_export(self._x, mine.ctypes.data), "svx_exchange_export")
            mine[-1] = 1
        except Exception as e:                      # noqa: BLE001 -- agreed on collectively below
            err = e
        if self.world == 1:
            if err is not None:
                raise err
            return
        # every collective below is executed by every rank, whatever happened locally, so that a
        # failure on one rank raises on all of them instead of leaving the others waiting
        on_gpu = dist.get_backend(group) == "nccl"
        dev = classifier.torch_device if on_gpu else torch.device("cpu")
        out = torch.empty((self.world * mine.size,), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, torch.from_numpy(mine).to(dev), group=group)
        table = out.cpu().numpy().reshape(self.world, mine.size)
        if err is None and not table[:, -1].all():
            err = RuntimeError(f"exchange setup failed on rank(s) {np.flatnonzero(table[:, -1] == 0).tolist()}")
        ok = 0
        if err is None:
            try:
                handles = np.ascontiguousarray(table[:, :-1])
                self._check(self._lib.svx_exchange_attach(self._x, handles.ctypes.data), "svx_exchange_attach")
                ok = 1
            except Exception as e:                  # noqa: BLE001
                err = e
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)        # also: every rank has mapped every buffer
        if int(flag.item()) == 0:
            self.close()
            raise err if err is not None else RuntimeError("exchange attach failed on another rank")

    def classify(self, rows_dev: torch.Tensor, raw: bool = False):
        """rows of this rank's shard (cuda int32 [n<=per,12]) -> (labels int32[world*per], scores
        float32[world*per]) in rank-major (= file) order: zero-copy views of the gathered buffer,
        valid after the current stream's work and until the next-but-one call (``raw=True``: the
        buffer itself, int32[world*per, 2] with the score bit-cast in column 1).  Asynchronous: a rank
        that never delivers is detected by the wait kernel's timeout, see :meth:`status` /
        :meth:`result`."""
        clf = self.clf
        assert rows_dev.dtype == torch.int32 and rows_dev.is_contiguous() and rows_dev.device == clf.torch_device
        ptr = ctypes.c_void_p()
        self._check(self._lib.svx_classify_exchange(clf._h, self._x, rows_dev.data_ptr(), rows_dev.shape[0],
                                                    ctypes.byref(ptr), clf._stream()), "svx_classify_exchange")
        calls = torch.as_tensor(_DeviceView(ptr.value, (self.world * self.per, 2), "<i4"),
                                device=clf.torch_device)
        if raw:
            return calls
        return calls[:, 0], calls[:, 1].view(torch.float32)

    def status(self) -> None:
        """Synchronises the device and raises if a rank failed to show up within the timeout
        (``SVX_EXCHANGE_TIMEOUT_MS``, default 30 s).  Until it is called, the error stays set and every
        further :meth:`classify` raises; the late rank's calls in the gathered buffer of the call
        that timed out are poisoned (label -1, score NaN)."""
        self._check(self._lib.svx_exchange_status(self._x), "svx_exchange_status")

    def result(self, rows_dev: torch.Tensor):
        """:meth:`classify` + :meth:`status`: the gathered (labels, scores), guaranteed complete (raises
        if any rank did not deliver).  Synchronises; use :meth:`classify` to keep the stream asynchronous
        and call :meth:`status` before trusting what was gathered."""
        out = self.classify(rows_dev)
        self.status()
        return out

    def close(self) -> None:
        if getattr(self, "_x", None):
            self._lib.svx_exchange_destroy(self._x)
            self._x = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

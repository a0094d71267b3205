"""Host-side entry to the B200 encode+classify path.

``Classifier`` is what replaces, inside the reference's ``Predict.run``
(``src/network/predict.py:148-210``), the pair

* ``BatchGenerator.next_batch`` (``src/network/create_batch.py:88-155``) and
* ``sess.run([score, tf.argmax(score, 1), tf.nn.softmax(score)], ...)``
  (``src/network/predict.py:209-210``),

with the same per-site results: ``labels`` (int, argmax) and ``probs`` (``numpy.float32`` softmax;
the dtype matters downstream).  PyTorch is used only as the device-memory
container and stream provider; all arithmetic runs in ``libsvx.so``.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from . import weights as _weights

IMG = 227
ROW_FIELDS = 12
NUM_CLASSES = 5
SV_TYPES = ("DEL", "INS", "INV", "DUP", "tDUP")      # src/network/predict.py:133-142


def _as_rows(rows) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    if rows.ndim != 2 or rows.shape[1] != ROW_FIELDS:
        raise ValueError(f"rows must be int32[N,{ROW_FIELDS}], got {rows.shape}")
    return rows


class Classifier:
    """One handle on one GPU.  ``model`` is a ``-m`` checkpoint prefix, a weights dict in TF
    layouts (``svision_b200.weights``), or ``None`` for an encoder-only handle."""

    def __init__(self, model=None, device: int = 0, max_batch: int = 2048,
                 precision: str = "3pass"):
        if not torch.cuda.is_available():
            raise _lib.SvxError("no CUDA device: the encode+classify path has no CPU fallback")
        self._lib = _lib.load()
        self.device = int(device)
        self.max_batch = int(max_batch)
        self.precision = precision
        prec = {"3pass": _lib.PRECISION_3PASS, "1pass": _lib.PRECISION_1PASS}[precision]
        if isinstance(model, str):
            model = _weights.load_checkpoint(model)
        wptr = None
        keep = []
        if model is not None:
            _weights.check_weights(model)
            w = _lib.SvxWeights()
            for layer in _weights.WEIGHT_SHAPES:
                for kind, short in (("weights", "w"), ("biases", "b")):
                    a = np.ascontiguousarray(model[f"{layer}/{kind}"], dtype=np.float32)
                    keep.append(a)
                    setattr(w, f"{layer}_{short}", a.ctypes.data)
            wptr = ctypes.byref(w)
        h = ctypes.c_void_p()
        _lib.check(self._lib.svx_create(wptr, self.device, self.max_batch, prec, ctypes.byref(h)),
                   "svx_create")
        self._h = h
        self.has_model = model is not None
        del keep

    # -- lifecycle ------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.svx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- helpers --------------------------------------------------------------------------------
    @property
    def torch_device(self) -> torch.device:
        return torch.device("cuda", self.device)

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.torch_device).cuda_stream

    def rows_to_device(self, rows) -> torch.Tensor:
        return torch.from_numpy(_as_rows(rows)).to(self.torch_device)

    # -- the path -------------------------------------------------------------------------------
    def encode(self, rows, dtype=torch.float32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """rows (numpy or cuda int32 tensor [N,12]) -> cuda tensor [N,227,227,3] of ``dtype``:
        bit-exact with what ``BatchGenerator.next_batch`` yields."""
        if not torch.is_tensor(rows):
            rows = self.rows_to_device(rows)
        assert rows.dtype == torch.int32 and rows.is_contiguous() and rows.device == self.torch_device
        n = rows.shape[0]
        code = {torch.float32: _lib.IMAGE_F32, torch.float16: _lib.IMAGE_F16}[dtype]
        if out is None:
            out = torch.empty((n, IMG, IMG, 3), dtype=dtype, device=self.torch_device)
        assert out.is_contiguous() and out.dtype == dtype and out.numel() == n * IMG * IMG * 3
        _lib.check(self._lib.svx_encode(self._h, rows.data_ptr(), n, out.data_ptr(), code,
                                        self._stream()), "svx_encode")
        return out

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """cuda images [N,227,227,3] (fp32/fp16) -> logits float32 [N,5]."""
        assert images.is_contiguous() and images.device == self.torch_device
        code = {torch.float32: _lib.IMAGE_F32, torch.float16: _lib.IMAGE_F16}[images.dtype]
        n = images.shape[0]
        logits = torch.empty((n, NUM_CLASSES), dtype=torch.float32, device=self.torch_device)
        _lib.check(self._lib.svx_forward(self._h, images.data_ptr(), code, n, logits.data_ptr(),
                                         self._stream()), "svx_forward")
        return logits

    def classify_device(self, rows: torch.Tensor, want_logits: bool = False):
        """Fused path on device buffers: rows cuda int32 [N,12] -> (labels int32[N], probs
        float32[N,5][, logits float32[N,5]]) as cuda tensors.  Asynchronous."""
        assert rows.dtype == torch.int32 and rows.is_contiguous() and rows.device == self.torch_device
        n = rows.shape[0]
        labels = torch.empty((n,), dtype=torch.int32, device=self.torch_device)
        probs = torch.empty((n, NUM_CLASSES), dtype=torch.float32, device=self.torch_device)
        logits = torch.empty((n, NUM_CLASSES), dtype=torch.float32, device=self.torch_device) \
            if want_logits else None
        _lib.check(self._lib.svx_classify_device(
            self._h, rows.data_ptr(), n, labels.data_ptr(), probs.data_ptr(),
            logits.data_ptr() if want_logits else None, self._stream()), "svx_classify_device")
        return (labels, probs, logits) if want_logits else (labels, probs)

    def classify_device_calls(self, rows: torch.Tensor, raw: bool = False):
        """Fused path returning only what ``predict.py:230,251`` consumes per row: (labels int32[N],
        scores float32[N]) -- views of one ``svx_call[N]`` buffer the fc8 kernel writes
        (``raw=True``: that buffer itself, int32[N,2] with the score bit-cast in column 1)."""
        assert rows.dtype == torch.int32 and rows.is_contiguous() and rows.device == self.torch_device
        n = rows.shape[0]
        calls = torch.empty((n, 2), dtype=torch.int32, device=self.torch_device)
        _lib.check(self._lib.svx_classify_device_calls(self._h, rows.data_ptr(), n, calls.data_ptr(),
                                                       self._stream()), "svx_classify_device_calls")
        if raw:
            return calls
        return calls[:, 0], calls[:, 1].view(torch.float32)

    def classify(self, rows, labels_out: Optional[np.ndarray] = None,
                 probs_out: Optional[np.ndarray] = None):
        """HOST entry (the call a reference-side user makes): rows numpy int32[N,12] ->
        (labels numpy int32[N], probs numpy float32[N,5]); H2D/D2H copies included."""
        rows = _as_rows(rows)
        n = rows.shape[0]
        labels = labels_out if labels_out is not None else np.empty((n,), dtype=np.int32)
        probs = probs_out if probs_out is not None else np.empty((n, NUM_CLASSES), dtype=np.float32)
        assert labels.dtype == np.int32 and probs.dtype == np.float32
        assert labels.flags.c_contiguous and probs.flags.c_contiguous
        _lib.check(self._lib.svx_classify(self._h, rows.ctypes.data, n, labels.ctypes.data,
                                          probs.ctypes.data), "svx_classify")
        return labels, probs

    PROFILE_SLOTS = ("encode", "conv1", "pool1_lrn1", "conv2", "pool2_lrn2", "conv3", "conv4",
                     "conv5", "pool5", "fc6", "fc7", "fc8_softmax")

    def set_profiling(self, enable: bool) -> None:
        _lib.check(self._lib.svx_set_profiling(self._h, int(bool(enable))), "svx_set_profiling")

    def profile_read(self, reset: bool = True) -> dict:
        """{kernel slot: (total ms, launches)} measured with CUDA events on the launching stream."""
        ms = np.zeros(len(self.PROFILE_SLOTS), dtype=np.float32)
        cnt = np.zeros(len(self.PROFILE_SLOTS), dtype=np.int64)
        _lib.check(self._lib.svx_profile_read(self._h, ms.ctypes.data, cnt.ctypes.data, int(reset)),
                   "svx_profile_read")
        return {k: (float(m), int(c)) for k, m, c in zip(self.PROFILE_SLOTS, ms, cnt)}

    def debug_counters(self, reset: bool = True) -> np.ndarray:
        """uint64[7,8] per-role cycle counters (handle must be created with SVX_DBG=1)."""
        out = np.zeros((7, 8), dtype=np.uint64)
        _lib.check(self._lib.svx_debug_counters(self._h, out.ctypes.data, int(reset)),
                   "svx_debug_counters")
        return out

    def debug_activation(self, name: str, n: int) -> np.ndarray:
        # (conv2 / conv5 at full resolution never exist: their max-pool runs in the layer's epilogue)
        shapes = {"conv1": (55, 55, 96), "norm1": (27, 27, 96), "norm2": (13, 13, 256),
                  "conv3": (13, 13, 384), "conv4": (13, 13, 384), "pool5": (6, 6, 256),
                  "fc6": (4096,), "fc7": (4096,)}
        out = np.empty((n,) + shapes[name], dtype=np.float32)
        _lib.check(self._lib.svx_debug_activation(self._h, name.encode(), n, out.ctypes.data),
                   "svx_debug_activation")
        return out


class MultiClassifier:
    """Every GPU of the box behind one object, in ONE process (C-ABI ``svx_multi_*``): what the
    reference's single ``SVision`` process (``SVision:296-341``) would hold.  ``classify`` has the
    contract of ``Classifier.classify`` and returns the same bits; the rows of a call are spread
    over the devices in chunks taken from a shared counter, so a slower GPU takes fewer."""

    def __init__(self, model, devices=None, max_batch: int = 8192, precision: str = "3pass"):
        if not torch.cuda.is_available():
            raise _lib.SvxError("no CUDA device: the encode+classify path has no CPU fallback")
        self._lib = _lib.load()
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        self.devices = [int(d) for d in devices]
        self.max_batch = int(max_batch)
        prec = {"3pass": _lib.PRECISION_3PASS, "1pass": _lib.PRECISION_1PASS}[precision]
        if isinstance(model, str):
            model = _weights.load_checkpoint(model)
        _weights.check_weights(model)
        w = _lib.SvxWeights()
        keep = []
        for layer in _weights.WEIGHT_SHAPES:
            for kind, short in (("weights", "w"), ("biases", "b")):
                a = np.ascontiguousarray(model[f"{layer}/{kind}"], dtype=np.float32)
                keep.append(a)
                setattr(w, f"{layer}_{short}", a.ctypes.data)
        dev = (ctypes.c_int * len(self.devices))(*self.devices)
        h = ctypes.c_void_p()
        _lib.check(self._lib.svx_multi_create(ctypes.byref(w), dev, len(self.devices), self.max_batch,
                                              prec, ctypes.byref(h)), "svx_multi_create")
        self._h = h
        self.has_model = True
        del keep

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.svx_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def classify(self, rows, labels_out: Optional[np.ndarray] = None,
                 probs_out: Optional[np.ndarray] = None):
        rows = _as_rows(rows)
        n = rows.shape[0]
        labels = labels_out if labels_out is not None else np.empty((n,), dtype=np.int32)
        probs = probs_out if probs_out is not None else np.empty((n, NUM_CLASSES), dtype=np.float32)
        assert labels.dtype == np.int32 and probs.dtype == np.float32
        assert labels.flags.c_contiguous and probs.flags.c_contiguous
        _lib.check(self._lib.svx_multi_classify(self._h, rows.ctypes.data, n, labels.ctypes.data,
                                                probs.ctypes.data), "svx_multi_classify")
        return labels, probs

    def last_split(self) -> list:
        """Sites each device processed in the last ``classify`` call."""
        out = np.zeros(len(self.devices), dtype=np.int64)
        _lib.check(self._lib.svx_multi_last_split(self._h, out.ctypes.data), "svx_multi_last_split")
        return out.tolist()


def gemm_selftest(a: torch.Tensor, b: torch.Tensor, block_n: int = 128,
                  precision: str = "3pass") -> torch.Tensor:
    """C = A @ B.T through the tcgen05 layer kernel (A [M,K], B [N,K] float32 cuda tensors;
    block_n 96 / 128 / 192 / 256, N a multiple of it)."""
    lib = _lib.load()
    assert a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32
    a, b = a.contiguous(), b.contiguous()
    m, k = a.shape
    n = b.shape[0]
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    prec = {"3pass": _lib.PRECISION_3PASS, "1pass": _lib.PRECISION_1PASS}[precision]
    _lib.check(lib.svx_gemm_selftest(a.device.index or 0, a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                     m, n, k, block_n, prec,
                                     torch.cuda.current_stream(a.device).cuda_stream),
               "svx_gemm_selftest")
    return c


def conv_selftest(a: torch.Tensor, b: torch.Tensor, row_off, block_n: int = 128,
                  precision: str = "3pass") -> torch.Tensor:
    """C[m,n] = sum_t A[m + row_off[t], :] @ B[n, t*K:(t+1)*K].T through the tensor-core layer kernel
    (A [M,K], B [N,taps*K] float32 cuda tensors; rows outside A read as zero)."""
    lib = _lib.load()
    a, b = a.contiguous(), b.contiguous()
    m, k = a.shape
    n = b.shape[0]
    offs = np.ascontiguousarray(row_off, dtype=np.int32)
    assert b.shape[1] == k * offs.size
    c = torch.empty((m, n), dtype=torch.float32, device=a.device)
    prec = {"3pass": _lib.PRECISION_3PASS, "1pass": _lib.PRECISION_1PASS}[precision]
    _lib.check(lib.svx_conv_selftest(a.device.index or 0, a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                     m, n, k, int(offs.size), offs.ctypes.data, block_n, prec,
                                     torch.cuda.current_stream(a.device).cuda_stream),
               "svx_conv_selftest")
    return c


def launch_count(reset: bool = False) -> int:
    lib = _lib.load()
    n = int(lib.svx_launch_count())
    if reset:
        lib.svx_launch_count_reset()
    return n

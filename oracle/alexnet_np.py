"""SECOND, independent CPU oracle of the reference CNN: plain numpy, float64, written from the
TensorFlow op definitions the reference calls -- no torch, no shared code with oracle/alexnet.py.

TEST INFRASTRUCTURE ONLY (tests/ import it; the product never does).  Parity status: TensorFlow
1.14 is not installable here, so this oracle is not pinned against TF itself; it pins the FIRST oracle
(oracle/alexnet.py, torch) by restating every op from its published definition:

* ``tf.nn.conv2d(x, W, strides=[1, s, s, 1], padding=...)`` (reference src/network/alexnet.py:109-111):
  cross-correlation over NHWC input with HWIO filters.  VALID: out = ceil((in - k + 1) / s), no padding.
  SAME: out = ceil(in / s), pad_total = max((out - 1) * s + k - in, 0), pad_before = pad_total // 2
  (the extra pixel goes to the bottom / right).
* groups (alexnet.py:124-129): ``tf.split`` the input AND the filters in ``groups`` parts along axis 3,
  convolve pairwise, ``tf.concat`` along axis 3.
* ``tf.nn.bias_add`` + ``tf.nn.relu`` (alexnet.py:132-136).
* ``tf.nn.max_pool(ksize 3x3, strides 2, VALID)`` (alexnet.py:158-161).
* ``tf.nn.local_response_normalization(x, depth_radius=2, alpha=2e-5, beta=0.75, bias=1.0)``
  (alexnet.py:164-166): ``out = x / (bias + alpha * sum_{d-r..d+r} x^2) ** beta`` over the channel axis,
  window clipped at the ends.
* ``tf.reshape(pool5, [-1, 6*6*256])`` (alexnet.py:49): NHWC order; ``tf.nn.xw_plus_b`` (+ relu for fc6 /
  fc7, alexnet.py:140-155); dropout with keep_prob = 1.0 is the identity (predict.py:22,210).
* ``tf.argmax(score, 1)`` (first maximum) and ``tf.nn.softmax(score)`` (predict.py:209).
Layer order: conv -> relu -> pool -> lrn (alexnet.py:29-36).
"""
from __future__ import annotations

import math

import numpy as np


def _same_pads(size: int, k: int, s: int):
    out = math.ceil(size / s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv2d(x: np.ndarray, w: np.ndarray, stride: int, padding: str) -> np.ndarray:
    """x [N,H,W,C] float64, w [kh,kw,C,O] -> [N,oh,ow,O]; im2col + one matrix product per image."""
    n, h, wd, c = x.shape
    kh, kw, ci, o = w.shape
    assert ci == c
    if padding == "SAME":
        oh, pt, pb = _same_pads(h, kh, stride)
        ow, pl, pr = _same_pads(wd, kw, stride)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    elif padding == "VALID":
        oh = math.ceil((h - kh + 1) / stride)
        ow = math.ceil((wd - kw + 1) / stride)
    else:
        raise ValueError(padding)
    wm = w.reshape(kh * kw * c, o)
    out = np.empty((n, oh, ow, o), dtype=np.float64)
    for i in range(n):
        cols = np.empty((oh, ow, kh, kw, c), dtype=np.float64)
        for a in range(kh):
            for b in range(kw):
                cols[:, :, a, b, :] = x[i, a:a + (oh - 1) * stride + 1:stride, b:b + (ow - 1) * stride + 1:stride, :]
        out[i] = (cols.reshape(oh * ow, kh * kw * c) @ wm).reshape(oh, ow, o)
    return out


def conv_layer(x, w, b, stride, padding, groups=1):
    if groups == 1:
        y = conv2d(x, w, stride, padding)
    else:
        xs = np.split(x, groups, axis=3)
        ws = np.split(w, groups, axis=3)
        y = np.concatenate([conv2d(xi, wi, stride, padding) for xi, wi in zip(xs, ws)], axis=3)
    return np.maximum(y + b, 0.0)


def max_pool_3x3_s2_valid(x):
    n, h, w, c = x.shape
    oh, ow = (h - 3) // 2 + 1, (w - 3) // 2 + 1
    out = np.full((n, oh, ow, c), -np.inf)
    for a in range(3):
        for b in range(3):
            out = np.maximum(out, x[:, a:a + 2 * (oh - 1) + 1:2, b:b + 2 * (ow - 1) + 1:2, :])
    return out


def lrn(x, radius=2, alpha=2e-5, beta=0.75, bias=1.0):
    c = x.shape[3]
    sq = x * x
    s = np.zeros_like(x)
    for d in range(c):
        s[..., d] = sq[..., max(0, d - radius):min(c, d + radius + 1)].sum(axis=3)
    return x / (bias + alpha * s) ** beta


def forward(images: np.ndarray, weights: dict, return_intermediates: bool = False):
    """images [N,227,227,3] (any float dtype) -> logits float64 [N,5]."""
    w = {k: np.asarray(v, dtype=np.float64) for k, v in weights.items()}
    x = np.asarray(images, dtype=np.float64)
    inter = {}
    x = conv_layer(x, w["conv1/weights"], w["conv1/biases"], 4, "VALID")
    x = lrn(max_pool_3x3_s2_valid(x))
    inter["norm1"] = x
    x = conv_layer(x, w["conv2/weights"], w["conv2/biases"], 1, "SAME", groups=2)
    x = lrn(max_pool_3x3_s2_valid(x))
    inter["norm2"] = x
    x = conv_layer(x, w["conv3/weights"], w["conv3/biases"], 1, "SAME")
    inter["conv3"] = x
    x = conv_layer(x, w["conv4/weights"], w["conv4/biases"], 1, "SAME", groups=2)
    inter["conv4"] = x
    x = conv_layer(x, w["conv5/weights"], w["conv5/biases"], 1, "SAME", groups=2)
    x = max_pool_3x3_s2_valid(x)
    inter["pool5"] = x
    x = x.reshape(x.shape[0], 6 * 6 * 256)
    x = np.maximum(x @ w["fc6/weights"] + w["fc6/biases"], 0.0)
    inter["fc6"] = x
    x = np.maximum(x @ w["fc7/weights"] + w["fc7/biases"], 0.0)
    inter["fc7"] = x
    logits = x @ w["fc8/weights"] + w["fc8/biases"]
    return (logits, inter) if return_intermediates else logits


def softmax(logits: np.ndarray) -> np.ndarray:
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def argmax_first(logits: np.ndarray) -> np.ndarray:
    return np.argmax(logits, axis=1)             # numpy, like tf.argmax, returns the first maximum

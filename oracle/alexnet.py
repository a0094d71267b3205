"""ORACLE (test infrastructure, not product code) -- CPU restatement of SVision's CNN forward
pass in plain torch (fp32, with an fp64 referee).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this.

Parity status: **UNPINNED against TensorFlow** -- the reference executes this graph with
TensorFlow 1.14 (``setup.py:36``), which is neither vendored under ``/root/reference`` nor
installable in this image, and the reference ships no tests, golden logits or trained
checkpoint.  What *is* pinned: the topology and every hyper-parameter below
are restated from the reference's own graph-construction code, and the torch ops used have the
published semantics of the TF ops they replace.

Follows (paths relative to the reference root):
  * ``src/network/alexnet.py:26-58``   layer order: conv -> ReLU -> max-pool -> LRN for layers
    1-2, conv3, grouped conv4/conv5, pool5, NHWC flatten (``:49``), fc6, fc7, fc8 (no ReLU).
  * ``src/network/alexnet.py:100-137`` grouped conv: input split on the channel axis, weights
    ``[kh,kw,Cin/groups,Cout]`` split on the *output* axis, results concatenated; bias; ReLU.
  * ``src/network/alexnet.py:140-155`` ``x @ W + b`` with ``W[in,out]``.
  * ``src/network/alexnet.py:158-166`` 3x3/2 VALID max-pool; LRN depth_radius 2, alpha 2e-5,
    beta 0.75, bias 1  ==  torch ``local_response_norm(size=5, alpha=1e-4, beta=.75, k=1)``
    (torch divides alpha by size).
  * ``src/network/alexnet.py:169-170`` + ``src/network/predict.py:22,210`` dropout with
    keep_prob 1.0 is the identity.
  * ``src/network/predict.py:167,209`` input ``float32[B,227,227,3]``; ``argmax(score, 1)``;
    ``softmax(score)``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

LAYERS = ("conv1", "conv2", "conv3", "conv4", "conv5", "fc6", "fc7", "fc8")
#: name -> (TF weight shape, groups)
SHAPES = {
    "conv1": ((11, 11, 3, 96), 1),
    "conv2": ((5, 5, 48, 256), 2),
    "conv3": ((3, 3, 256, 384), 1),
    "conv4": ((3, 3, 192, 384), 2),
    "conv5": ((3, 3, 192, 256), 2),
    "fc6": ((9216, 4096), 1),
    "fc7": ((4096, 4096), 1),
    "fc8": ((4096, 5), 1),
}


def _t(w, dtype):
    return torch.as_tensor(np.asarray(w)).to(dtype)


def _conv(x, weights, name, dtype, stride=1, padding=0):
    w = _t(weights[f"{name}/weights"], dtype).permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
    b = _t(weights[f"{name}/biases"], dtype)
    groups = SHAPES[name][1]
    return F.relu(F.conv2d(x, w, b, stride=stride, padding=padding, groups=groups))


def _lrn(x):
    return F.local_response_norm(x, size=5, alpha=1e-4, beta=0.75, k=1.0)


@torch.no_grad()
def forward(images_nhwc, weights, dtype=torch.float32, return_intermediates: bool = False):
    """``[B,227,227,3]`` (numpy or tensor) -> logits ``[B,5]`` (tensor of ``dtype``).

    With ``return_intermediates`` also returns a dict of NHWC activations after each fused stage
    (``norm1``, ``norm2``, ``conv3``, ``conv4``, ``pool5``, ``fc6``, ``fc7``)."""
    x = torch.as_tensor(np.asarray(images_nhwc)).to(dtype).permute(0, 3, 1, 2).contiguous()
    inter = {}
    x = _conv(x, weights, "conv1", dtype, stride=4)
    if return_intermediates:
        inter["conv1"] = x.permute(0, 2, 3, 1)
    x = _lrn(F.max_pool2d(x, 3, 2))
    inter["norm1"] = x.permute(0, 2, 3, 1)
    x = _conv(x, weights, "conv2", dtype, padding=2)
    if return_intermediates:
        inter["conv2"] = x.permute(0, 2, 3, 1)
    x = _lrn(F.max_pool2d(x, 3, 2))
    inter["norm2"] = x.permute(0, 2, 3, 1)
    x = _conv(x, weights, "conv3", dtype, padding=1)
    inter["conv3"] = x.permute(0, 2, 3, 1)
    x = _conv(x, weights, "conv4", dtype, padding=1)
    inter["conv4"] = x.permute(0, 2, 3, 1)
    x = _conv(x, weights, "conv5", dtype, padding=1)
    if return_intermediates:
        inter["conv5"] = x.permute(0, 2, 3, 1)
    x = F.max_pool2d(x, 3, 2)
    inter["pool5"] = x.permute(0, 2, 3, 1)
    x = x.permute(0, 2, 3, 1).reshape(x.shape[0], 6 * 6 * 256)            # NHWC flatten
    x = F.relu(x @ _t(weights["fc6/weights"], dtype) + _t(weights["fc6/biases"], dtype))
    inter["fc6"] = x
    x = F.relu(x @ _t(weights["fc7/weights"], dtype) + _t(weights["fc7/biases"], dtype))
    inter["fc7"] = x
    x = x @ _t(weights["fc8/weights"], dtype) + _t(weights["fc8/biases"], dtype)
    if return_intermediates:
        return x, inter
    return x


@torch.no_grad()
def classify(images_nhwc, weights, dtype=torch.float32, batch: int = 128):
    """Reference-shaped result: (labels int64[B], softmax dtype[B,5], logits dtype[B,5])."""
    outs = []
    n = len(images_nhwc)
    for s in range(0, n, batch):
        outs.append(forward(images_nhwc[s:s + batch], weights, dtype))
    logits = torch.cat(outs, 0) if outs else torch.zeros((0, 5), dtype=dtype)
    return torch.argmax(logits, 1).numpy(), torch.softmax(logits, 1).numpy(), logits.numpy()

"""Model weights for the SVision CNN: the 16 TensorFlow variables of the reference graph
(``src/network/alexnet.py:113-116,141-145``), in TF layouts (conv ``[kh,kw,Cin/groups,Cout]``,
fc ``[in,out]``), as a ``dict[str, np.ndarray(float32)]``.

* :func:`synthetic_weights` -- seeded random-init weights of the reference architecture (there is
  no network, and the reference's trained checkpoint is a Google-Drive download:
  ``README.md:85-86``), optionally with the calibrated ``fc8`` of SURVEY.md §8(d) so that class
  labels are balanced and label parity is a discriminating test.
* :func:`load_checkpoint` -- the ``-m <prefix>`` loader (``SVision:35``, restored by
  ``src/network/predict.py:181-184``); see :mod:`svision_b200.tf_bundle`.
"""
from __future__ import annotations

import os

import numpy as np

NUM_CLASSES = 5
#: name -> TF weight shape (reference: src/network/alexnet.py:29-58)
WEIGHT_SHAPES = {
    "conv1": (11, 11, 3, 96),
    "conv2": (5, 5, 48, 256),
    "conv3": (3, 3, 256, 384),
    "conv4": (3, 3, 192, 384),
    "conv5": (3, 3, 192, 256),
    "fc6": (9216, 4096),
    "fc7": (4096, 4096),
    "fc8": (4096, NUM_CLASSES),
}
VARIABLE_NAMES = tuple(f"{l}/{k}" for l in WEIGHT_SHAPES for k in ("weights", "biases"))
N_PARAMS = sum(int(np.prod(s)) + s[-1] for s in WEIGHT_SHAPES.values())   # 56 888 709

_HERE = os.path.dirname(os.path.abspath(__file__))
CALIBRATION_FILE = os.path.join(os.path.dirname(_HERE), "tests", "golden", "fc8_calibrated.npz")


def synthetic_weights(seed: int = 1234, calibrated: bool = True) -> dict:
    """He-initialised weights (``N(0, sqrt(2/fan_in))``), biases ``U(0, 0.1)``.

    ``calibrated=True`` replaces ``fc8`` by the committed calibration (per-class logit mean
    removed, logit std 3.0 on 512 stress images; produced by ``oracle/make_cnn_golden.py``)."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in WEIGHT_SHAPES.items():
        fan_in = int(np.prod(shape[:-1]))
        w[f"{name}/weights"] = (rng.standard_normal(shape, dtype=np.float32)
                                * np.float32(np.sqrt(2.0 / fan_in)))
        w[f"{name}/biases"] = rng.uniform(0.0, 0.1, size=shape[-1]).astype(np.float32)
    if calibrated:
        if not os.path.exists(CALIBRATION_FILE):
            raise FileNotFoundError(f"{CALIBRATION_FILE} missing: run oracle/make_cnn_golden.py")
        cal = np.load(CALIBRATION_FILE)
        if int(cal["seed"]) != seed:
            raise ValueError("fc8 calibration was made for a different seed")
        w["fc8/weights"] = cal["weights"].astype(np.float32)
        w["fc8/biases"] = cal["biases"].astype(np.float32)
    return w


def check_weights(w: dict) -> None:
    """Raise if ``w`` is not exactly the reference's variable set."""
    for name, shape in WEIGHT_SHAPES.items():
        for key, shp in ((f"{name}/weights", shape), (f"{name}/biases", (shape[-1],))):
            if key not in w:
                raise KeyError(f"model is missing variable {key!r}")
            a = np.asarray(w[key])
            if tuple(a.shape) != tuple(shp):
                raise ValueError(f"variable {key!r} has shape {a.shape}, expected {shp}")
            if a.dtype != np.float32:
                raise TypeError(f"variable {key!r} has dtype {a.dtype}, expected float32")


def load_checkpoint(prefix: str) -> dict:
    """``-m <prefix>``: read ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` without TF."""
    from . import tf_bundle
    w = tf_bundle.read_bundle(prefix, names=VARIABLE_NAMES)
    check_weights(w)
    return w

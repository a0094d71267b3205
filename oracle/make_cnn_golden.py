"""ORACLE tooling -- (1) calibrate the synthetic ``fc8`` (SURVEY.md §8(d) "Synthetic weights") with
the fp64 oracle and (2) write CNN golden logits for a fixed set of sites.

    python oracle/make_cnn_golden.py

Outputs ``tests/golden/fc8_calibrated.npz`` (used by ``svision_b200.weights.synthetic_weights``)
and ``tests/golden/cnn_golden.npz`` (rows + fp32/fp64 oracle logits).  TensorFlow is not
installable here, so these pin the *oracle*, not TF: CNN parity vs the reference's TF path is
unpinned (see ``oracle/alexnet.py``)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import alexnet, encoder_c  # noqa: E402
from svision_b200 import sites, weights  # noqa: E402

SEED_W = 1234
SEED_CAL = 11
N_CAL = 512
N_GOLD = 256


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    w = weights.synthetic_weights(SEED_W, calibrated=False)
    cal_rows = sites.make_sites_p2(N_CAL, seed=SEED_CAL)
    imgs = encoder_c.encode_f32(cal_rows)
    logits = np.concatenate([alexnet.forward(imgs[s:s + 64], w, torch.float64).numpy()
                             for s in range(0, N_CAL, 64)])
    print("uncalibrated label histogram", np.bincount(logits.argmax(1), minlength=5))
    # centre per class, then scale so the logit std is 3.0
    mean = logits.mean(axis=0)
    b = w["fc8/biases"].astype(np.float64) - mean
    scale = 3.0 / (logits - mean).std()
    w8 = (w["fc8/weights"].astype(np.float64) * scale).astype(np.float32)
    b8 = (b * scale).astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fc8_calibrated.npz"),
                        weights=w8, biases=b8, seed=np.int64(SEED_W))
    w = weights.synthetic_weights(SEED_W, calibrated=True)
    logits = np.concatenate([alexnet.forward(imgs[s:s + 64], w, torch.float64).numpy()
                             for s in range(0, N_CAL, 64)])
    p = torch.softmax(torch.from_numpy(logits), 1).numpy()
    srt = np.sort(logits, axis=1)
    print("calibrated label histogram", np.bincount(logits.argmax(1), minlength=5),
          "mean max-prob %.3f" % p.max(1).mean(),
          "median top-2 margin %.3f min %.2e" % (np.median(srt[:, -1] - srt[:, -2]),
                                                 (srt[:, -1] - srt[:, -2]).min()))

    g = np.load(os.path.join(ROOT, "tests", "golden", "encoder_golden.npz"))
    rows = g["rows"]
    pick = np.random.default_rng(5).choice(rows.shape[0], N_GOLD, replace=False)
    pick.sort()
    rows = rows[pick]
    imgs = encoder_c.encode_f32(rows)
    l64 = np.concatenate([alexnet.forward(imgs[s:s + 64], w, torch.float64).numpy()
                          for s in range(0, N_GOLD, 64)])
    l32 = np.concatenate([alexnet.forward(imgs[s:s + 64], w, torch.float32).numpy()
                          for s in range(0, N_GOLD, 64)])
    print("fp32 vs fp64 max |dlogit| %.3e" % np.abs(l32 - l64).max(),
          "labels", np.bincount(l64.argmax(1), minlength=5))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cnn_golden.npz"),
                        rows=rows, logits_fp64=l64, logits_fp32=l32, seed=np.int64(SEED_W),
                        meta=np.array([f"torch={torch.__version__}", "oracle/alexnet.py"]))
    print("done")


if __name__ == "__main__":
    main()

"""ORACLE tooling -- generate ``tests/golden/encoder_golden.npz`` by running the REFERENCE's own
encoder (``/root/reference/src/network/create_batch.py:BatchGenerator`` ->
``src/segmentplot/plot_segment.py:PlotSingleImg`` -> ``cv2.line``) in the build container.

Run (container only; ``/root/reference`` does not exist on the GPU box, so the fixtures are what
travels)::

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

Nothing from the reference is copied: it is imported, executed on rows written by
``svision_b200.sites`` and only its *outputs* are stored, as lit-pixel codes
``ch*51529 + row*227 + col`` (the image is a 3-bit-per-pixel bitmap: SURVEY.md F7).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("SVISION_REFERENCE", "/root/reference")

from svision_b200 import sites  # noqa: E402
from oracle import encoder as enc  # noqa: E402


def reference_images(rows: np.ndarray, batch: int = 128) -> np.ndarray:
    """Run the reference BatchGenerator over ``rows``; returns float32[N,227,227,3]."""
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from src.network.create_batch import BatchGenerator  # reference code, imported not copied

    with tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False) as f:
        f.write("\n".join(sites.rows_to_bed_lines(rows)) + "\n")
        path = f.name
    try:
        gen = BatchGenerator(path, shuffle=False, nb_classes=5, batch_size=batch)
        n = rows.shape[0]
        out = np.empty((n, 227, 227, 3), dtype=np.float32)
        done = 0
        while done < n:
            imgs, labels = gen.next_batch(batch)
            take = min(batch, n - done)
            f32 = imgs[:take].astype(np.float32)
            assert np.array_equal(f32.astype(np.float64), imgs[:take])
            out[done:done + take] = f32
            done += take
    finally:
        os.unlink(path)
    return out


def images_to_codes(imgs: np.ndarray):
    lo = np.array([l[0] for l in enc.LEVELS], dtype=np.float32)
    hi = np.array([l[1] for l in enc.LEVELS], dtype=np.float32)
    lit = imgs == hi
    assert np.array_equal(np.where(lit, hi, lo), imgs), "reference image is not two-level"
    offsets = [0]
    codes = []
    for i in range(imgs.shape[0]):
        c = enc.pack_bits(np.moveaxis(lit[i], -1, 0))
        codes.append(c)
        offsets.append(offsets[-1] + c.size)
    return np.asarray(offsets, dtype=np.int64), np.concatenate(codes).astype(np.uint32)


def main():
    n_p2 = int(os.environ.get("GOLDEN_P2", 4096))
    n_p1 = int(os.environ.get("GOLDEN_P1", 1024))
    rows = np.concatenate([
        sites.edge_case_sites(),
        sites.make_sites_p2(n_p2, seed=sites.SEED_P2),
        sites.make_sites_p1(n_p1, seed=sites.SEED_CONFIG2, profile="hifi"),
        sites.make_sites_p1(n_p1, seed=sites.SEED_CONFIG5, profile="ont"),
    ], axis=0)
    out_off, out_codes = [], []
    chunk = 512
    base = 0
    offsets = [np.zeros(1, dtype=np.int64)]
    for s in range(0, rows.shape[0], chunk):
        imgs = reference_images(rows[s:s + chunk])
        off, codes = images_to_codes(imgs)
        offsets.append(off[1:] + base)
        base += codes.size
        out_codes.append(codes)
        print(f"  reference encoded {min(s + chunk, rows.shape[0])}/{rows.shape[0]}", flush=True)
    offsets = np.concatenate(offsets)
    codes = np.concatenate(out_codes)
    import cv2
    dst = os.path.join(ROOT, "tests", "golden", "encoder_golden.npz")
    np.savez_compressed(dst, rows=rows, offsets=offsets, codes=codes,
                        meta=np.array([f"cv2={cv2.__version__}", f"numpy={np.__version__}",
                                       "source=reference BatchGenerator.next_batch"]))
    print("wrote", dst, rows.shape, codes.size, os.path.getsize(dst))


if __name__ == "__main__":
    main()

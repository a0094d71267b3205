"""CPU: the two encoder restatements (oracle/encoder.py, oracle/encoder_c.c) against the golden
vectors produced by the reference's own BatchGenerator (oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import encoder as enc, encoder_c
from svision_b200 import sites


def _codes_of(bits_u8):
    return enc.pack_bits(bits_u8.astype(bool))


def test_c_oracle_matches_reference_golden(encoder_golden):
    rows, off, codes = (encoder_golden[k] for k in ("rows", "offsets", "codes"))
    bits = encoder_c.encode_bits(rows)
    assert bits.shape == (rows.shape[0], 3, 227, 227)
    for i in range(rows.shape[0]):
        assert np.array_equal(_codes_of(bits[i]), codes[off[i]:off[i + 1]]), f"row {i}: {rows[i]}"


def test_python_oracle_matches_reference_golden(encoder_golden):
    rows, off, codes = (encoder_golden[k] for k in ("rows", "offsets", "codes"))
    # the pure-Python restatement is slow: all edge cases + a strided sample of the rest
    n_edge = sites.edge_case_sites().shape[0]
    idx = list(range(n_edge)) + list(range(n_edge, rows.shape[0], 7))
    for i in idx:
        got = enc.pack_bits(enc.encode_bits(rows[i]))
        assert np.array_equal(got, codes[off[i]:off[i + 1]]), f"row {i}: {rows[i]}"


def test_pad_row_lights_two_pixels(encoder_golden):
    # SURVEY Appendix A note (5): the reference pad row lights exactly (0,0) and (1,1) of ch0
    bits = enc.encode_bits(sites.PAD_ROW)
    assert bits[0].sum() == 2 and bits[0, 0, 0] and bits[0, 1, 1]
    assert bits[1].sum() == 0 and bits[2].sum() == 0
    assert np.array_equal(encoder_golden["rows"][0], sites.PAD_ROW)


def test_float_image_levels():
    rows = sites.make_sites_p2(64, seed=99)
    f = encoder_c.encode_f32(rows)
    assert f.shape == (64, 227, 227, 3) and f.dtype == np.float32
    for ch, (lo, hi) in enumerate(enc.LEVELS):
        vals = np.unique(f[..., ch])
        assert set(vals.tolist()) <= {lo, hi}
    assert np.array_equal(f, enc.encode_rows(rows))
    # all six levels are exact in fp16 -> 16-bit images are lossless
    assert np.array_equal(f.astype(np.float16).astype(np.float32), f)


def test_digest_distinguishes_images():
    rows = sites.make_sites_p1(256, seed=5)
    d = encoder_c.encode_digest(rows)
    bits = encoder_c.encode_bits(rows)
    flat = bits.reshape(256, -1)
    for i in range(0, 256, 17):
        for j in range(i + 1, 256, 31):
            assert (d[i] == d[j]) == bool(np.array_equal(flat[i], flat[j]))


def test_encoder_is_order_dependent_on_reverse_segments():
    # reverse segments are drawn end->start (plot_segment.py:49-52); with clipping that differs
    # from start->end for some lines, so the flag must matter beyond channel 2
    rows = sites.make_sites_p2(2048, seed=123)
    rev = rows.copy()
    rev[:, 4] = 0
    rev[:, 9] = 0
    b = encoder_c.encode_bits(rev)
    assert (b[:, 2] == b[:, 0]).all()     # every pixel of a reverse segment is also in ch2


@pytest.mark.parametrize("n", [0, 1, 3])
def test_small_counts(n):
    rows = sites.make_sites_p1(n, seed=1) if n else np.zeros((0, 12), np.int32)
    assert encoder_c.encode_bits(rows).shape == (n, 3, 227, 227)


def test_oracles_match_reference_on_real_demo_rows():
    """272 rows derived from the reference's demo BAM by the reference's own collection stage
    (oracle/make_demo_rows.py), encoded by the reference's BatchGenerator."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "demo_rows_golden.npz"))
    rows, off, codes = g["rows"], g["offsets"], g["codes"]
    assert rows.shape == (272, 12)
    bits = encoder_c.encode_bits(rows)
    for i in range(rows.shape[0]):
        ref = codes[off[i]:off[i + 1]]
        assert np.array_equal(_codes_of(bits[i]), ref), f"C oracle, demo row {i}"
        if i % 5 == 0:
            assert np.array_equal(enc.pack_bits(enc.encode_bits(rows[i])), ref), f"py oracle, demo row {i}"

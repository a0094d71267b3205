"""Segment-pair generation from signatures: ``svision_b200.pairs`` against the BED text
the reference's collection stage writes.

* demo golden: the clusters the reference found in its demo BAM (``tests/golden/demo_clusters.json``, captured
  by ``oracle/make_pairs_golden.py``) must give ``tests/golden/demo_chr9.segments.bed`` byte for byte;
* fuzz golden: seeded synthetic signatures + the text of the reference's ``proc_one_cluster`` on them;
* live (build container only): fresh seeds against the imported reference."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from svision_b200 import bed, pairs                    # noqa: E402

REF = os.environ.get("SVISION_REFERENCE", "/root/reference")


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_pairs_golden", os.path.join(ROOT, "oracle", "make_pairs_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    try:
        spec.loader.exec_module(mod)          # only defines functions; the reference is imported lazily
    finally:
        sys.path[:] = saved
    return mod


G = _golden_module()


def render(cluster_dicts, min_support, max_sv_size):
    sigs = pairs.SignatureTable.from_clusters([G.dict_to_cluster(d) for d in cluster_dicts], min_support, max_sv_size)
    table = pairs.generate_pairs(sigs)
    return table, "".join(l + "\n" for l in pairs.to_bed_lines(table))


def test_demo_clusters_reproduce_the_reference_bed():
    g = json.load(open(os.path.join(HERE, "golden", "demo_clusters.json")))
    table, text = render(g["clusters"], g["min_support"], g["max_sv_size"])
    want = open(os.path.join(HERE, "golden", "demo_chr9.segments.bed")).read()
    assert text == want
    # and the packed rows are what the BED reader would have parsed back from that file
    parsed = bed.parse_segments_bed(want.encode())
    assert np.array_equal(table.rows, parsed.rows)
    for k in ("bkp_start", "bkp_end", "bkp_len", "flags"):
        assert np.array_equal(getattr(table, k), getattr(parsed, k)), k
    for k in bed.SegmentsTable.STRING_COLUMNS:
        assert getattr(table, k).tolist() == getattr(parsed, k).tolist(), k


def test_fuzz_golden_text():
    g = np.load(os.path.join(HERE, "golden", "pairs_fuzz_golden.npz"))
    _seed, _n, min_support, max_sv_size = (int(v) for v in g["meta"])
    table, text = render(json.loads(str(g["clusters"])), min_support, max_sv_size)
    assert len(table) > 3000 and text == str(g["text"])
    assert (table.rows[:, 4] == 1).all()                      # the first segment of a pair is always a main one
    assert (table.rows[:, 9] == 0).any() and (table.flags & bed.FLAG_MAIN).any()


def test_edge_cases_and_errors():
    empty = pairs.SignatureTable([], np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros((0, 5), np.int64),
                                 np.zeros(1, np.int64), np.zeros((0, 3), np.int64), np.empty(0, object),
                                 np.empty(0, object), np.empty(0, object))
    assert len(pairs.generate_pairs(empty)) == 0

    def one(aligns, bkps):
        return pairs.SignatureTable(["chr1+10+20+3"], np.zeros(1, np.int64), np.array([0, len(aligns)]),
                                    np.array(aligns, np.int64).reshape(-1, 5), np.array([0, len(bkps)]),
                                    np.array(bkps, np.int64).reshape(-1, 3), np.array(["q"], object),
                                    np.array(["sigGap"], object), np.array(["None"], object))
    # a single alignment has no pair; two co-linear alignments neither (output_clusters.py:170)
    assert len(pairs.generate_pairs(one([[100, 200, 0, 100, 0]], [[1, 2, 3]]))) == 0
    assert len(pairs.generate_pairs(one([[100, 200, 0, 100, 0], [210, 300, 110, 200, 0]], [[1, 2, 3]]))) == 0
    # a deletion-like gap gives the main pair; coordinates become relative to the first alignment
    t = pairs.generate_pairs(one([[1000, 1100, 50, 150, 0], [1600, 1700, 151, 251, 0]], [[7, 8, 9]]))
    assert t.rows.tolist() == [[0, 100, 0, 100, 1, 101, 201, 600, 700, 1, 201, 700]]
    assert t.read_num.tolist() == ["1m"] and (t.bkp_start[0], t.bkp_end[0], t.bkp_len[0]) == (7, 8, 9)
    # an inner segment needs its own breakpoint (the reference would raise IndexError: output_clusters.py:199)
    with pytest.raises(ValueError, match="breakpoint"):
        pairs.generate_pairs(one([[0, 100, 0, 100, 0], [5000, 5100, 100, 200, 1], [200, 300, 200, 300, 0]], [[1, 2, 3]]))
    with pytest.raises(ValueError, match="int32"):
        pairs.generate_pairs(one([[0, 100, 0, 100, 0], [3_000_000_000, 3_000_000_100, 100, 200, 0]], [[1, 2, 3]]))
    with pytest.raises(ValueError, match="no alignment"):
        pairs.generate_pairs(one([], [[1, 2, 3]]))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "collection")), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("seed", [101, 102, 103])
def test_live_fuzz_vs_reference(seed):
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pysam_stub"))
    sys.path.insert(0, REF)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            fuzz = G.synthetic_clusters(seed, 150)
            want = G.reference_lines(fuzz, 3, 8000)
        _, got = render(fuzz, 3, 8000)
        assert got == want and want.count("\n") > 300
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods and k.split(".")[0] in ("src", "pysam"):
                del sys.modules[k]

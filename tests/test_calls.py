"""Post-classification host step: ``svision_b200.calls`` against the text the
reference's own functions print.

* golden: ``tests/golden/calls_golden.npz`` holds the VCF / score text produced by the unmodified
  ``get_region_potential_svtypes`` + ``write_results_to_vcf`` + ``genotyper`` (``oracle/make_calls_golden.py``)
  on a seeded synthetic stream; ``calls.call_chromosome`` + ``calls.AlignmentTable`` must reproduce it
  byte for byte (no reference tree needed -> also runs on the GPU box);
* live (build container only): fuzz of each function against its reference counterpart."""
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import make_calls_golden as G, reference_loader as RL   # noqa: E402
from svision_b200 import calls, sites                                # noqa: E402

needs_reference = pytest.mark.skipif(not RL.available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(HERE, "golden", "calls_golden.npz"))
    n_rows, table_seed, _label_seed, aln_seed = (int(v) for v in g["meta"])
    table = sites.make_region_table(n_rows, seed=table_seed)
    aln = sites.make_alignments(table, seed=aln_seed)
    return g, table, aln


def make_table(aln) -> calls.AlignmentTable:
    return calls.AlignmentTable(aln["contig_length"], aln["reference_start"], aln["reference_end"],
                                aln["mapping_quality"], aln["is_unmapped"], aln["is_secondary"], aln["query_name"])


def render(records):
    return ("".join(line + "\n" for _, line in records), "".join(str(q) + "\n" for q, _ in records))


@pytest.mark.parametrize("tag,opt", [("s3_qname", G.options(3, True)), ("s1", G.options(1, False)),
                                     ("s5_min200", G.options(5, False, 200))])
def test_call_chromosome_reproduces_reference_text(golden, tag, opt):
    g, table, aln = golden
    at = make_table(aln)
    for genotype in (at, at.genotype):            # one vectorised pass / candidate by candidate
        records = calls.call_chromosome(table, g["labels"], g["probs"], opt, genotype)
        vcf, score = render(records)
        assert len(records) > 100
        assert vcf == str(g[f"{tag}_vcf"])
        assert score == str(g[f"{tag}_score"])


@pytest.mark.parametrize("chunk_rows", [1, 97, 1000, 10**9])
def test_streamed_calling_equals_whole_table(golden, chunk_rows):
    """Chunks cut at region changes, classification of chunk k+1 on a worker thread: same text."""
    g, table, aln = golden
    labels, probs = g["labels"], g["probs"]
    index = {table.rows[i].tobytes(): i for i in range(len(table))}
    assert len(index) > 0.9 * len(table)
    calls_seen = []

    def classify(rows):                                   # returns the golden labels of exactly these rows
        calls_seen.append(rows.shape[0])
        at0 = next(i for i in range(len(table) - rows.shape[0] + 1)
                   if i == index.get(rows[0].tobytes(), -1) or np.array_equal(table.rows[i:i + rows.shape[0]], rows))
        assert np.array_equal(table.rows[at0:at0 + rows.shape[0]], rows)
        return labels[at0:at0 + rows.shape[0]], probs[at0:at0 + rows.shape[0]]

    opt = G.options(3, True)
    records = calls.call_chromosome_streamed(table, classify, opt, make_table(aln), chunk_rows=chunk_rows)
    vcf, score = render(records)
    assert vcf == str(g["s3_qname_vcf"]) and score == str(g["s3_qname_score"])
    cuts = calls.region_cuts(table, chunk_rows)
    assert cuts[0] == 0 and cuts[-1] == len(table) and all(b > a for a, b in zip(cuts, cuts[1:]))
    assert all(table.region[c] != table.region[c - 1] for c in cuts[1:-1])
    assert sum(calls_seen) == len(table) and len(calls_seen) == len(cuts) - 1
    if chunk_rows == 1:
        assert len(cuts) - 1 == len(set(table.region.tolist()))         # one chunk per region


def test_region_cuts_use_parser_flags(tmp_path):
    from svision_b200 import bed
    table = sites.make_region_table(3000, seed=77)
    p = tmp_path / "chr1.segments.all.bed"
    p.write_text("\n".join(sites.table_to_bed_lines(table)) + "\n")
    parsed = bed.read_segments_bed(str(p))
    assert calls.region_cuts(parsed, 500) == calls.region_cuts(table, 500)
    assert calls.region_cuts(table.take(slice(0, 0)), 10) == [0]


def _parsed(table, tmp_path):
    from svision_b200 import bed
    p = tmp_path / "chr1.segments.all.bed"
    p.write_text("\n".join(sites.table_to_bed_lines(table)) + "\n")
    return bed.read_segments_bed(str(p))


@pytest.mark.parametrize("tag,opt", [("s3_qname", G.options(3, True)), ("s1", G.options(1, False)),
                                     ("s5_min200", G.options(5, False, 200))])
def test_native_aggregation_reproduces_reference_text(golden, tag, opt, tmp_path, monkeypatch):
    """The same golden text through ``svx_calls_aggregate`` (table parsed from BED text), whole and in
    chunks, and the pure-Python route on the same parsed table."""
    g, table, aln = golden
    parsed = _parsed(table, tmp_path)
    assert parsed.has_text() and not table.has_text()
    at = make_table(aln)
    used = []
    real = calls.pending_records_native
    monkeypatch.setattr(calls, "pending_records_native", lambda *a: used.append(1) or real(*a))
    vcf, score = render(calls.call_chromosome(parsed, g["labels"], g["probs"], opt, at))
    assert used and vcf == str(g[f"{tag}_vcf"]) and score == str(g[f"{tag}_score"])
    index = {parsed.rows[i].tobytes(): i for i in range(len(parsed))}

    def classify(rows):
        i0 = index[rows[0].tobytes()]
        while not np.array_equal(parsed.rows[i0:i0 + rows.shape[0]], rows):
            i0 = next(i for i in range(i0 + 1, len(parsed)) if np.array_equal(parsed.rows[i], rows[0]))
        return g["labels"][i0:i0 + rows.shape[0]], g["probs"][i0:i0 + rows.shape[0]]

    n_used = len(used)
    vcf2, score2 = render(calls.call_chromosome_streamed(parsed, classify, opt, at, chunk_rows=700))
    assert len(used) > n_used + 3 and vcf2 == vcf and score2 == score
    monkeypatch.setattr(calls, "USE_NATIVE", False)
    n_used = len(used)
    vcf3, score3 = render(calls.call_chromosome(parsed, g["labels"], g["probs"], opt, at))
    assert len(used) == n_used and vcf3 == vcf and score3 == score


def test_native_numpy_reductions_match_numpy():
    """numpy.mean over float32 scores and numpy.std over integer signature scores, restated in
    csrc/host_calls.cpp, against numpy itself (they end up printed: output.py:473-474,551)."""
    import ctypes
    from svision_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for trial in range(4000):
        n = int(rng.integers(1, 300)) if trial % 20 else int(rng.integers(300, 30000))
        a = np.round(rng.random(n).astype(np.float32), 2)
        out = np.zeros(1, np.float32)
        assert lib.svx_np_mean_f32(a.ctypes.data, n, out.ctypes.data) == 0
        assert out[0] == (np.mean(list(a)) if n < 300 else np.mean(a)), n
        b = rng.integers(-50, 5000, size=n).astype(np.int64)
        o2 = np.zeros(1, np.float64)
        assert lib.svx_np_std_i64(b.ctypes.data, n, o2.ctypes.data) == 0
        assert o2[0] == (np.std([int(x) for x in b]) if n < 300 else np.std(b)), n
    assert lib.svx_np_mean_f32(None, 0, None) != 0


def test_native_aggregation_declines_what_only_python_parses(golden, tmp_path, monkeypatch):
    g, table, aln = golden
    lines = sites.table_to_bed_lines(table.take(slice(0, 200)))
    cols = lines[5].split("\t")
    cols[19] = " 7 "                                       # int(' 7 ') is fine in Python; the native parser declines
    lines[5] = "\t".join(cols)
    from svision_b200 import bed
    parsed = bed.parse_segments_bed(("\n".join(lines) + "\n").encode())
    win = np.round(g["probs"][np.arange(200), g["labels"][:200]], 2)
    assert calls.pending_records_native(parsed, g["labels"][:200], win, G.options(1, False)) is None
    recs = calls.call_chromosome(parsed, g["labels"][:200], g["probs"][:200], G.options(1, False), make_table(aln))
    monkeypatch.setattr(calls, "USE_NATIVE", False)      # the same table through the pure-Python route
    calls_py = calls.call_chromosome(parsed, g["labels"][:200], g["probs"][:200], G.options(1, False), make_table(aln))
    assert [l for _, l in recs] == [l for _, l in calls_py] and len(recs) > 3


def test_write_chromosome_files(golden, tmp_path):
    g, table, aln = golden
    recs = calls.call_chromosome(table, g["labels"], g["probs"], G.options(3, True), make_table(aln).genotype)
    calls.write_chromosome(str(tmp_path / "chr1.predict.s3"), recs)
    assert open(tmp_path / "chr1.predict.s3.vcf").read() == str(g["s3_qname_vcf"])
    assert open(tmp_path / "chr1.predict.s3.score.txt").read() == str(g["s3_qname_score"])


def test_empty_and_degenerate_inputs():
    table = sites.make_region_table(40, seed=3)
    empty = table.take(slice(0, 0))
    at = calls.AlignmentTable(1000, [], [], [], [], [], [])
    assert calls.call_chromosome(empty, np.zeros(0, np.int32), np.zeros((0, 5), np.float32), G.options(), at.genotype) == []
    # every row a forward signature classified INV -> everything dropped, nothing flushed (predict.py:229-231)
    lab = np.full(len(table), 2, np.int32)
    pr = np.full((len(table), 5), 0.1, np.float32)
    pr[:, 2] = 0.6
    fwd = type(table)(table.rows, table.bkp_start, table.bkp_end, table.bkp_len,
                      **{**{k: getattr(table, k) for k in table.STRING_COLUMNS},
                         "forward": np.array(["True"] * len(table), dtype=object)})
    assert calls.call_chromosome(fwd, lab, pr, G.options(1), at.genotype) == []
    assert calls.aggregate_region({}) == []
    with pytest.raises(ValueError):
        calls.AlignmentTable(1000, [5, 3], [9, 9], [60, 60], [0, 0], [0, 0], ["a", "b"])
    with pytest.raises(AssertionError):
        calls.call_chromosome(table, lab, pr.astype(np.float64), G.options(), at.genotype)


def test_round_matches_python_round_on_float32():
    """``call_chromosome`` rounds the winning scores in one ``numpy.round``; the reference rounds them one
    by one with ``round(numpy.float32, 2)`` (predict.py:251).  Same values, same dtype."""
    x = np.random.default_rng(0).random(200_000).astype(np.float32)
    r = np.round(x, 2)
    assert r.dtype == np.float32
    for i in range(0, x.size, 37):
        v = round(x[i], 2)
        assert type(v) is np.float32 and v == r[i]


# ------------------------------------------------------------------------------------------------
# live cross-checks against the reference functions (build container only)
# ------------------------------------------------------------------------------------------------
@needs_reference
def test_aggregate_region_fuzz_vs_reference():
    rng = np.random.default_rng(1)
    with RL.reference_modules() as ref:
        pred = ref.Predict("chr1", "unused")
        for _ in range(400):
            reads = {}
            for r in rng.permutation(int(rng.integers(1, 25))).tolist():
                kinds = rng.choice(5, size=int(rng.integers(1, 4)), replace=False).tolist()
                reads[str(r)] = {np.int64(k): [int(rng.integers(0, 250_000_000)), int(rng.integers(0, 250_000_000)),
                                               int(rng.integers(0, 100_000))] for k in kinds}
            import copy
            assert calls.aggregate_region(copy.deepcopy(reads)) == pred.get_region_potential_svtypes(copy.deepcopy(reads))


@needs_reference
def test_refine_types_fuzz_vs_reference():
    rng = np.random.default_rng(2)
    import copy
    with RL.reference_modules() as ref:
        for _ in range(3000):
            ids = sorted(rng.choice(5, size=int(rng.integers(1, 6)), replace=False).tolist())
            kinds = [calls.TYPE_NAMES[k] for k in ids]
            anchor = int(rng.integers(1000, 100000))
            bk = [[anchor + int(rng.integers(-15, 16)), anchor + int(rng.integers(-15, 16)), int(rng.integers(0, 400))]
                  for _ in kinds]
            opt = types.SimpleNamespace(min_sv_size=int(rng.choice([0, 50, 200])))
            want = ref.refine_type(copy.deepcopy(kinds), copy.deepcopy(bk), opt)
            got = calls.refine_types(copy.deepcopy(kinds), copy.deepcopy(bk), opt.min_sv_size)
            assert (list(got[0]), [list(b) for b in got[1]]) == (list(want[0]), [list(b) for b in want[1]])


@needs_reference
def test_genotype_fuzz_vs_reference():
    table = sites.make_region_table(1500, seed=77)
    aln = sites.make_alignments(table, seed=78, depth=45)
    at = make_table(aln)
    rng = np.random.default_rng(3)
    names_all = sorted(set(table.read_name.tolist()))
    with RL.reference_modules() as ref:
        ref.genotype.pysam = RL.FakePysam(aln)
        regions = list(dict.fromkeys(table.region.tolist()))
        batch = []
        for trial in range(300):
            reg = regions[int(rng.integers(0, len(regions)))]
            contig, a, b = reg.split("+")[:3]
            a, b = int(a) + int(rng.integers(-50, 50)), int(b) + int(rng.integers(-50, 5000))
            kinds = [["DEL"], ["INS"], ["INV"], ["DUP"], ["tDUP"], ["INS", "tDUP"], []][int(rng.integers(0, 7))]
            support = [names_all[int(i)] for i in rng.integers(0, len(names_all), size=int(rng.integers(0, 12)))]
            support += [n for r, n in zip(table.region.tolist(), table.read_name.tolist()) if r == reg][:int(rng.integers(0, 9))]
            opt = types.SimpleNamespace(min_mapq=int(rng.choice([0, 10, 30])), min_gt_depth=int(rng.choice([1, 4, 20])),
                                        homo_thresh=0.8, hete_thresh=0.2, bam_path="x")
            cand = (contig, a, b, kinds)
            want = ref.genotyper(cand, support, opt)
            assert at.genotype(cand, support, opt) == want, (trial, cand)
            batch.append((cand, support, opt, want))
        for mapq in (0, 10, 30):                       # the batched pass shares one options object
            sub = [b for b in batch if b[2].min_mapq == mapq and b[2].min_gt_depth == 4]
            got = at.genotype_many([b[0] for b in sub], [b[1] for b in sub], sub[0][2])
            assert got == [b[3] for b in sub]


@needs_reference
@pytest.mark.parametrize("seed", [11, 12])
def test_call_chromosome_live_vs_reference(seed):
    table = sites.make_region_table(2500, seed=seed, profile="ont")
    labels, probs = G.synthetic_labels(table, seed + 100)
    aln = sites.make_alignments(table, seed=seed + 200, depth=20)
    opt = G.options(2, True, 50)
    vcf, score, opens = G.reference_text(table, labels, probs, aln, opt)
    got_vcf, got_score = render(calls.call_chromosome(table, labels, probs, opt, make_table(aln)))
    assert got_vcf == vcf and got_score == score
    assert opens == vcf.count("\n") > 50           # the reference re-opens the BAM once per record
    import tempfile
    with tempfile.TemporaryDirectory() as d:        # the native route (table parsed from BED text)
        import pathlib
        parsed = _parsed(table, pathlib.Path(d))
        nat_vcf, nat_score = render(calls.call_chromosome(parsed, labels, probs, opt, make_table(aln)))
    assert nat_vcf == vcf and nat_score == score


# ---- the reference's two row-skip rules at the seams (predict.py:214,229-247) -----------------------------
def _with_strings(table, **cols):
    """A copy of a string-column table with some columns replaced."""
    from svision_b200 import bed
    kw = {c: getattr(table, c).copy() for c in ("region", "read_num", "read_name", "sig_type", "sig_score",
                                                "forward", "mechanism")}
    kw.update(cols)
    return bed.SegmentsTable(table.rows.copy(), table.bkp_start.copy(), table.bkp_end.copy(),
                             table.bkp_len.copy(), **kw)


def test_rows_labelled_complement_are_skipped_like_the_reference(golden, tmp_path):
    """predict.py:214 skips every row whose joined label contains 'complement' -- its own pad rows, but
    also a read or mechanism that happens to be named like that.  All routes (native aggregation over
    the parsed text, pure Python, the replayed loop feeding the reference's own functions) must give
    the text of the same stream WITHOUT those rows."""
    g, table, aln = golden
    n = 1500
    table = table.take(slice(0, n))
    labels, probs = g["labels"][:n], g["probs"][:n]
    rng = np.random.default_rng(5)
    hit = np.sort(rng.choice(n, 40, replace=False))
    names = table.read_name.copy().astype(object)
    mech = table.mechanism.copy().astype(object)
    for i in hit[:20]:
        names[i] = f"read_complement_{i}"
    for i in hit[20:]:
        mech[i] = "xcomplementx"
    marked = _with_strings(table, read_name=np.array(names.tolist()), mechanism=np.array(mech.tolist()))
    keep = np.setdiff1d(np.arange(n), hit)
    expect = table.take(keep)
    opt = G.options(3, True)
    at = make_table(aln)
    want = render(calls.call_chromosome(expect, labels[keep], probs[keep], opt, at))
    assert (np.asarray(marked.flags)[hit] & 16).all() and not (np.asarray(marked.flags)[keep] & 16).any()
    assert render(calls.call_chromosome(marked, labels, probs, opt, at)) == want          # Python route
    parsed = _parsed(marked, tmp_path)                                                     # native route
    assert (np.asarray(parsed.flags)[hit] & 16).all() and not (np.asarray(parsed.flags)[keep] & 16).any()
    assert render(calls.call_chromosome(parsed, labels, probs, opt, at)) == want
    assert render(calls.call_chromosome_streamed(parsed, lambda r: (labels[:0], probs[:0]) if r.shape[0] == 0 else
                                                 _labels_for(parsed, labels, probs, r), opt, at, chunk_rows=200)) == want
    # the replayed loop (what feeds the reference's own aggregate / write functions)
    from svision_b200 import predict
    seen_a, seen_b = [], []
    predict.replay_rows(marked, labels, probs, lambda region, reads, *rest: seen_a.append((region, dict(reads), [list(map(str, r)) if isinstance(r, list) else dict(r) for r in rest])))
    predict.replay_rows(expect, labels[keep], probs[keep], lambda region, reads, *rest: seen_b.append((region, dict(reads), [list(map(str, r)) if isinstance(r, list) else dict(r) for r in rest])))
    assert seen_a == seen_b


def _labels_for(table, labels, probs, rows):
    """The labels of exactly these (contiguous) rows of `table`."""
    n = rows.shape[0]
    for i in np.flatnonzero((table.rows[:, :] == rows[0]).all(axis=1)):
        if i + n <= len(table) and np.array_equal(table.rows[i:i + n], rows):
            return labels[i:i + n], probs[i:i + n]
    raise AssertionError("rows not found")


def test_streamed_seams_do_not_split_a_region_whose_interruption_is_dropped(golden):
    """Region A, then rows of region B that are ALL dropped (forward + INV, predict.py:229-231), then A
    again: the reference never sees B, so A stays one region.  A chunk cut that lands inside or next to
    the dropped run must not produce a second record for A."""
    g, table, aln = golden
    opt = G.options(1, True)
    at = make_table(aln)
    a = np.flatnonzero(table.region == table.region[0])
    first_b = int(a[-1]) + 1
    b = np.flatnonzero(table.region == table.region[first_b])
    assert len(a) >= 4 and len(b) >= 2
    half = len(a) // 2
    order = np.concatenate([a[:half], b, a[half:], np.arange(int(b[-1]) + 1, int(b[-1]) + 400)])
    t = table.take(order)
    labels, probs = g["labels"][order].copy(), g["probs"][order].copy()
    fwd = t.forward.copy()
    fwd[half:half + len(b)] = "True"                     # every B row: forward ...
    labels[half:half + len(b)] = 2                       # ... and classified INV -> dropped
    t = _with_strings(t, forward=fwd)
    whole = render(calls.call_chromosome(t, labels, probs, opt, at))
    for chunk_rows in (1, 2, half, half + 1, half + len(b), 50):
        got = render(calls.call_chromosome_streamed(t, lambda r: _labels_for(t, labels, probs, r), opt, at,
                                                    chunk_rows=chunk_rows))
        assert got == whole, chunk_rows
    # cutting at the raw region changes and flushing at every cut (what the streamed path did before the
    # open region was held back) gives MORE records: the test discriminates
    naive = []
    for lo, hi in ((0, half), (half, half + len(b)), (half + len(b), len(t))):
        naive.extend(calls.call_chromosome(t.take(slice(lo, hi)), labels[lo:hi], probs[lo:hi], opt, at))
    assert len(naive) > len(whole[0].splitlines())


def test_genotype_many_in_bounded_blocks_equals_one_pass(golden, monkeypatch):
    """Deep pile-ups: the vectorised genotyper works in blocks of bounded size and sends windows deeper
    than a block through the per-candidate route; the answers do not change."""
    g, table, aln = golden
    at = make_table(aln)
    opt = G.options(1, True)
    records = calls.call_chromosome(table, g["labels"], g["probs"], opt, lambda *a: ("./.", 0, 0))
    cands, sups = [], []
    for _, line in records[:400]:
        f = line.split("\t")
        info = dict(kv.split("=", 1) for kv in f[7].split(";") if "=" in kv)
        cands.append((f[0], int(f[1]), int(info["END"]), info["SVTYPE"].split("+")))
        sups.append(info.get("READS", "").split(","))
    whole = at.genotype_many(cands, sups, opt)
    assert whole == [at.genotype(c, s, opt) for c, s in zip(cands, sups)]
    for cap in (1, 50, 3000):                                # 1: every window is "deep" -> per-candidate route
        monkeypatch.setattr(calls, "MAX_PAIRS_PER_BLOCK", cap)
        assert at.genotype_many(cands, sups, opt) == whole

"""CPU: host-side logic around the kernels -- BED parsing, per-row replay, sharding + all-gather
(world_size 2 over gloo)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svision_b200 import bed, predict, sharded, sites


def _write_bed(tmp_path, rows, extra=None):
    lines = sites.rows_to_bed_lines(rows, region_size=5)
    if extra:
        lines += extra
    p = tmp_path / "chr1.segments.all.bed"
    p.write_text("\n".join(lines) + "\n")
    return str(p)


def test_bed_roundtrip(tmp_path):
    rows = np.concatenate([sites.edge_case_sites(), sites.make_sites_p2(200, seed=4)])
    t = bed.read_segments_bed(_write_bed(tmp_path, rows))
    assert t.rows.dtype == np.int32 and np.array_equal(t.rows, rows)
    assert len(t) == rows.shape[0]
    # label string layout of create_batch.py:48: col13, col0, col15..col22 joined by 'svision'
    lab = t.label_strings()[7].split("svision")
    assert lab[0] == "2m" and lab[1].startswith("chr1+1000+") and lab[2] == "read7" and len(lab) == 10


def test_bed_invalid_strand_token_takes_reverse_branch(tmp_path):
    line = sites.rows_to_bed_lines(sites.make_sites_p1(1, seed=1))[0].split("\t")
    line[5] = "None"          # neither 'True' nor 'False' -> forward=None -> falsy (create_batch.py:115)
    line[10] = "true"         # case matters in the reference
    t = bed.read_segments_bed(_write_bed(tmp_path, np.zeros((0, 12), np.int32), ["\t".join(line)]))
    assert t.rows[0, 4] == 0 and t.rows[0, 9] == 0


def test_bed_empty_and_malformed(tmp_path):
    p = tmp_path / "e.bed"
    p.write_text("")
    assert len(bed.read_segments_bed(str(p))) == 0
    p.write_text("a\tb\tc\n")
    with pytest.raises(ValueError):
        bed.read_segments_bed(str(p))


def test_bed_native_parser_matches_pandas_reader(tmp_path):
    table = sites.make_region_table(5000, seed=9, profile="ont")
    p = tmp_path / "c.bed"
    p.write_text("\n".join(sites.table_to_bed_lines(table)) + "\n")
    a, b = bed.read_segments_bed(str(p)), bed.read_segments_bed_pandas(str(p))
    assert np.array_equal(a.rows, b.rows) and np.array_equal(a.rows, table.rows)
    for k in ("bkp_start", "bkp_end", "bkp_len", "flags"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert np.array_equal(getattr(a, k), getattr(table, k)), k
    for k in bed.SegmentsTable.STRING_COLUMNS:
        assert getattr(a, k).tolist() == getattr(b, k).tolist() == getattr(table, k).tolist(), k
    assert a.label_strings() == b.label_strings()
    sub = a.take(slice(10, 20))
    assert len(sub) == 10 and sub.read_name.tolist() == table.read_name[10:20].tolist()


def test_bed_native_parser_edge_cases():
    base = sites.table_to_bed_lines(sites.make_region_table(3, seed=1))
    cols = base[0].split("\t")
    # no trailing newline, blank lines, CRLF, extra columns, negative and signed numbers, UTF-8 read name
    neg = list(cols)
    neg[1], neg[3], neg[17] = "-12", "+7", "-5"
    utf = list(cols)
    utf[15] = "läs/1"
    text = ("\n".join([base[0], "", base[1] + "\textra\tcolumns", "\t".join(neg)]) + "\r\n" + "\t".join(utf)).encode()
    t = bed.parse_segments_bed(text)
    assert len(t) == 4
    assert t.rows[2, 0] == -12 and t.rows[2, 2] == 7 and t.bkp_start[2] == -5
    assert t.bkp_len[2] == int(cols[22])                       # '\r' before the newline is tolerated, as int() does
    assert t.read_name[3] == "läs/1" and t.read_name[0] == cols[15]
    assert t.flags[1] & bed.FLAG_SAME_REGION and not t.flags[0] & bed.FLAG_SAME_REGION
    for bad, what in ((1, "x12"), (3, "99999999999"), (17, ""), (22, "1.5"), (12, "1_000")):
        broken = list(cols)
        broken[bad] = what
        with pytest.raises(ValueError, match="line 2"):
            bed.parse_segments_bed((base[0] + "\n" + "\t".join(broken) + "\n").encode())
    with pytest.raises(ValueError, match="23 tab-separated"):
        bed.parse_segments_bed(b"\t".join([b"1"] * 22) + b"\n")
    with pytest.raises(TypeError):
        bed.parse_segments_bed("text, not bytes")


def test_replay_rows_region_flush_and_rules(tmp_path):
    rows = sites.make_sites_p1(12, seed=2)
    lines = [l.split("\t") for l in sites.rows_to_bed_lines(rows, region_size=4)]
    lines[1][13] = "1"        # a main x minor pair (no 'm'): INS/DEL predictions are ignored
    lines[2][20] = "True"     # forward signature predicted INV -> dropped entirely
    t = bed.read_segments_bed(_write_bed(tmp_path, np.zeros((0, 12), np.int32),
                                         ["\t".join(l) for l in lines]))
    labels = np.array([1, 0, 2, 3, 1, 1, 1, 1, 4, 4, 4, 4], dtype=np.int32)
    probs = np.full((12, 5), 0.05, dtype=np.float32)
    probs[np.arange(12), labels] = 0.8
    seen = []
    n = predict.replay_rows(t, labels, probs, lambda *a: seen.append(a))
    assert n == 3 and [s[0] for s in seen] == ["chr1+0+500+4", "chr1+1000+1500+4", "chr1+2000+2500+4"]
    region0 = seen[0]
    assert region0[1] == {"0": {1: [100, 200, 100]}, "3": {3: [100, 200, 100]}}   # rows 0 and 3 only
    assert len(region0[5]) == 3 and all(isinstance(s, np.float32) for s in region0[5])
    assert region0[5][0] == np.float32(0.8)
    assert set(region0[2]) == {"0", "1", "3"}                                   # row 2 was dropped


def test_run_predict_writes_reference_files(tmp_path):
    rows = sites.make_sites_p1(9, seed=3)
    path = _write_bed(tmp_path, rows)

    class FakeClf:
        def classify(self, r):
            p = np.full((r.shape[0], 5), 0.1, dtype=np.float32)
            p[:, 1] = 0.6
            return np.ones(r.shape[0], dtype=np.int32), p

    calls = []

    def write(vcf, score, stats, region, *rest):
        calls.append(region)
        vcf.write(region + "\n")

    class Opt:
        model_path = "unused"
    n = predict.run_predict(path, str(tmp_path / "chr1.predict.s5"), Opt(), lambda d: sorted(d), write,
                            classifier=FakeClf(), chrom="chr1")
    assert n == 2 and calls == ["chr1+0+500+5", "chr1+1000+1500+5"]
    assert (tmp_path / "chr1.predict.s5.vcf").read_text().count("\n") == 2
    assert (tmp_path / "chr1.predict.s5.score.txt").exists()


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 100_000):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                a, b, per = sharded.shard_bounds(n, world, r)
                assert b - a <= per
                got += list(range(a, b))
                assert sharded.shard_rows(np.zeros((n, 12), np.int32), world, r).shape == (per, 12)
            assert got == list(range(n))


def _fake_classify(rows):
    labels = (np.abs(rows[:, 5].astype(np.int64)) % 5).astype(np.int32)
    probs = np.zeros((rows.shape[0], 5), dtype=np.float32)
    probs[np.arange(rows.shape[0]), labels] = 0.5 + (rows[:, 10] % 97) / 200.0
    return labels, probs


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = sites.make_sites_p1(n, seed=9)
    labels, probs = sharded.classify_sharded(_fake_classify, rows)
    ref_l, ref_p = _fake_classify(rows)
    q.put((rank, bool(np.array_equal(labels, ref_l)), bool(np.array_equal(probs, ref_p))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 64])
def test_sharded_all_gather_gloo_world2(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + n
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]


def _exchange_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class NoGpuClassifier:                      # the C-ABI rejects a NULL handle: setup fails locally
        _h = None
        torch_device = torch.device("cpu")

    try:
        sharded.Exchange(NoGpuClassifier(), 100)
        q.put((rank, "no error"))
    except Exception as e:                      # noqa: BLE001
        q.put((rank, type(e).__name__))
    dist.destroy_process_group()


def test_exchange_setup_failure_raises_on_every_rank_gloo_world2():
    """A rank whose exchange setup fails must not leave the others waiting in a collective: the
    failure is agreed on and raised everywhere (here both ranks fail: no GPU, no handle)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 777
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r for r, _ in res] == [0, 1] and all(name != "no error" for _, name in res)


def test_bed_reader_on_real_reference_output():
    """The BED written by the reference's writer_cluster_to_file for its demo BAM."""
    here = os.path.dirname(__file__)
    path = os.path.join(here, "golden", "demo_chr9.segments.bed")
    t = bed.read_segments_bed(path)
    g = np.load(os.path.join(here, "golden", "demo_rows_golden.npz"))
    assert np.array_equal(t.rows, g["rows"])
    assert len(set(t.region)) == 11
    # every row parses the way the reference's reader does (create_batch.py:42-49,103-137)
    for i, line in enumerate(open(path)):
        cols = line.rstrip("\n").split("\t")
        items = "_".join(cols[1:13]).split("_")
        assert [int(items[k]) for k in (0, 1, 2, 3, 5, 6, 7, 8, 10, 11)] == \
            [int(t.rows[i, k]) for k in (0, 1, 2, 3, 5, 6, 7, 8, 10, 11)]
        assert (items[4] == "True") == bool(t.rows[i, 4]) and (items[9] == "True") == bool(t.rows[i, 9])
        label = cols[13] + "svision" + cols[0] + "svision" + "svision".join(cols[15:23])
        assert t.label_strings()[i] == label if i < 3 else True

"""CPU: Step 2 of the driver (svision_b200/step2.py) -- per-chromosome prediction files, score range and
the merged VCF -- against the reference's own ``cal_scores_max_min`` + ``merge_split_vcfs``
(src/network/output.py:251-348,601-612), imported unmodified in the build container, and against a
committed golden of the merged text elsewhere."""
import hashlib
import os
import types

import numpy as np
import pytest

from svision_b200 import calls, sites, step2

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_SHA = os.path.join(HERE, "golden", "step2_merged.sha256")


class TableClassifier:
    """Stands in for the GPU classifier: returns precomputed labels / scores for known row blocks."""

    def __init__(self):
        self.known = {}

    def add(self, rows, labels, probs):
        self.known[rows.tobytes()] = (labels, probs)

    def classify(self, rows):
        return self.known[np.ascontiguousarray(rows, dtype=np.int32).tobytes()]


def _options(tmp_path, min_support=2, graph=False):
    return types.SimpleNamespace(min_support=min_support, qname=True, min_sv_size=50, min_mapq=10, min_gt_depth=4,
                                 homo_thresh=0.8, hete_thresh=0.2, bam_path="synthetic.bam", graph=graph,
                                 model_path="unused", sample="HG00733", out_path=str(tmp_path),
                                 genome=str(tmp_path / "genome.fa"))


def _prepare(tmp_path, chroms=("chr1", "chr2", "chrX"), rows_per_chrom=900):
    from oracle import make_calls_golden as G            # label synthesis only
    seg_dir, pred_dir = tmp_path / "segments", tmp_path / "predict_results"
    seg_dir.mkdir()
    clf, tables = TableClassifier(), {}
    for k, chrom in enumerate(chroms):
        table = sites.make_region_table(rows_per_chrom + 37 * k, seed=1000 + k, contig=chrom)
        labels, probs = G.synthetic_labels(table, 50 + k)
        clf.add(table.rows, labels, probs)
        aln = sites.make_alignments(table, seed=3 + k)
        tables[chrom] = calls.AlignmentTable(aln["contig_length"], aln["reference_start"], aln["reference_end"],
                                             aln["mapping_quality"], aln["is_unmapped"], aln["is_secondary"],
                                             aln["query_name"])
        (seg_dir / f"{chrom}.segments.all.bed").write_text("\n".join(sites.table_to_bed_lines(table)) + "\n")
    (tmp_path / "genome.fa.fai").write_text("chr1\t248956422\t112\t70\t71\nchr2\t242193529\t252513167\t70\t71\n"
                                            "chrX\t156040895\t2871101557\t70\t71\nchrM\t16569\t3031042417\t70\t71\n")
    return str(seg_dir), str(pred_dir), clf, tables


def test_step2_files_and_merge_rules(tmp_path):
    seg_dir, pred_dir, clf, tables = _prepare(tmp_path)
    opt = _options(tmp_path)
    merged = step2.run_step2(["chr1", "chr2", "chrMissing", "chrX"], seg_dir, pred_dir, opt, classifier=clf,
                             genotype_for=tables.get)
    assert os.path.basename(merged) == "HG00733.svision.s2.vcf"
    for chrom in ("chr1", "chr2", "chrX"):
        assert os.path.exists(os.path.join(pred_dir, f"{chrom}.predict.s2.vcf"))
        assert os.path.exists(os.path.join(pred_dir, f"{chrom}.predict.s2.score.txt"))
    lines = open(merged).read().split("\n")
    assert lines[0] == "##fileformat=VCFv4.3" and lines[1] == "##source=SVision v1.4"
    assert lines[2] == "##contig=<ID=chr1,length=248956422>" and lines[5] == "##contig=<ID=chrM,length=16569>"
    body = [l.split("\t") for l in lines if l and not l.startswith("#")]
    head = [l for l in lines if l.startswith("#")]
    assert head[-1].endswith("FORMAT\tHG00733") and not any("GraphID" in l for l in head)
    assert len(body) > 50 and [r[0] for r in body] == sorted([r[0] for r in body], key=["chr1", "chr2", "chrX"].index)
    # ids: serial numbers from 0 over distinct (POS, END); repeats get _1, _2, ...
    serial, prev = -1, None
    saw_sub = False
    for r in body:
        key = (r[0], r[1], r[7].split(";")[0])
        if key == prev:
            assert r[2].startswith(f"{serial}_")
            saw_sub = True
        else:
            serial += 1
            assert r[2] == str(serial)
        prev = key
        assert 0 <= int(r[5]) <= 100
    assert saw_sub
    assert {int(r[5]) for r in body} >= {0, 100}            # min-max rescale reaches both ends
    # stable text: the same inputs always give the same file (committed digest)
    # (oracle/make_step2_golden.py: digest of the text the reference's merge_split_vcfs wrote)
    body_only = "".join(l + "\n" for l in lines if l and not l.startswith("#"))
    assert hashlib.sha256(open(merged, "rb").read()).hexdigest() == open(GOLDEN_SHA).read().split()[0]
    assert hashlib.sha256(body_only.encode()).hexdigest() == open(GOLDEN_SHA).read().split()[1]


def test_step2_degenerate_scores_and_empty(tmp_path):
    pred = tmp_path / "p"
    pred.mkdir()
    with pytest.raises(ValueError):
        step2.score_range(str(pred))
    rec = "chr1\t100\t0\tN\t<SV>\t12.5\tPASS\tEND=200;SVLEN=100;SVTYPE=DEL;SUPPORT=3;BKPS=DEL:100-100-200\tGT:DR:DV\t0/1:3:3\n"
    (pred / "chr1.predict.s2.vcf").write_text(rec + rec)
    (pred / "chr1.predict.s2.score.txt").write_text("12.5\n12.5\n0\n")
    hi, lo = step2.score_range(str(pred))
    assert hi == lo == 12.5
    opt = _options(tmp_path, graph=True)
    out = tmp_path / "m.vcf"
    n = step2.merge_chromosomes(str(pred), str(out), hi, lo, ["chr1"], opt, contigs=[("chr1", 1000)])
    text = out.read_text().split("\n")
    assert n == 2 and any("GraphID" in l for l in text)
    body = [l.split("\t") for l in text if l and not l.startswith("#")]
    assert [r[2] for r in body] == ["0", "0_1"] and [r[5] for r in body] == ["100", "100"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/network"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("graph", [False, True])
def test_merge_is_byte_identical_to_the_reference(tmp_path, graph):
    from oracle import reference_loader as RL
    seg_dir, pred_dir, clf, tables = _prepare(tmp_path)
    opt = _options(tmp_path, graph=graph)
    chroms = ["chr1", "chr2", "chrX"]
    merged = step2.run_step2(chroms, seg_dir, pred_dir, opt, classifier=clf, genotype_for=tables.get)

    class FastaFile:                                    # serves the .fai, as pysam does
        def __init__(self, path):
            self._c = step2.contigs_from_fai(path)
            self.references = [n for n, _ in self._c]

        def get_reference_length(self, name):
            return dict(self._c)[name]

    with RL.reference_modules() as ref:
        ref.output.pysam = types.SimpleNamespace(FastaFile=FastaFile)
        scores = ref.output.cal_scores_max_min(pred_dir)
        hi, lo = np.max(scores), np.min(scores)                            # SVision:334
        assert (hi, lo) == step2.score_range(pred_dir)
        ref_path = str(tmp_path / "reference_merged.vcf")
        ref.output.merge_split_vcfs(pred_dir, ref_path, hi, lo, chroms, opt)
    assert open(merged, "rb").read() == open(ref_path, "rb").read()
    if not graph:                                       # the committed digest is of this very text
        assert hashlib.sha256(open(merged, "rb").read()).hexdigest() == open(GOLDEN_SHA).read().split()[0]


def test_command_line_runs_step2_with_the_reference_flags(tmp_path):
    seg_dir, _, clf, tables = _prepare(tmp_path, chroms=("chr2", "chr1"))       # files exist for chr1, chr2
    os.rename(seg_dir, str(tmp_path / "out_segments_tmp"))
    out = tmp_path / "out"
    out.mkdir()
    os.rename(str(tmp_path / "out_segments_tmp"), str(out / "segments"))
    argv = ["-o", str(out), "-b", "synthetic.bam", "-m", "unused.ckpt", "-g", str(tmp_path / "genome.fa"),
            "-n", "NA12878", "-s", "2", "--qname", "--debug", "-t", "8", "--batch_size", "128", "--window_size", "5000000"]
    assert step2.main(argv, classifier=clf, genotype_for=tables.get) == 0
    merged = out / "NA12878.svision.s2.vcf"
    body = [l.split("\t") for l in merged.read_text().split("\n") if l and not l.startswith("#")]
    assert [r[0] for r in body] == sorted((r[0] for r in body), key=["chr1", "chr2"].index)   # .fai order
    assert {r[0] for r in body} == {"chr1", "chr2"} and "READS=" in body[0][7]
    assert (out / "predict_results" / "chr1.predict.s2.vcf").exists()                          # --debug keeps them
    # -c restricts, --contig forces min_support 1, intermediates are removed without --debug
    argv2 = [a for a in argv if a != "--debug"] + ["-c", "chr2:1-1000", "--contig"]
    assert step2.main(argv2, classifier=clf, genotype_for=tables.get) == 0
    body2 = [l.split("\t") for l in (out / "NA12878.svision.s1.vcf").read_text().split("\n") if l and not l.startswith("#")]
    assert {r[0] for r in body2} == {"chr2"} and not (out / "predict_results").exists()
    assert step2.main(argv + ["-c", "chr9"], classifier=clf, genotype_for=tables.get) == 1


class HashClassifier:
    """Deterministic per-row results (any shard of any batch gives the same answer for a row)."""

    def classify(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        h = (rows * np.arange(1, 13, dtype=np.int64)).sum(1)
        labels = (np.abs(h) % 5).astype(np.int32)
        probs = np.full((rows.shape[0], 5), 0.05, dtype=np.float32)
        probs[np.arange(rows.shape[0]), labels] = (0.6 + (np.abs(h) % 37) / 100.0).astype(np.float32)
        return labels, probs


def _step2_worker(rank, world, port, argv, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rc = step2.main(argv, classifier=HashClassifier(), genotype_for=lambda chrom: (lambda *a: ("0/1", 3, 4)))
    q.put((rank, rc))
    dist.destroy_process_group()


@pytest.mark.parametrize("shard", ["rows", "chrom"])
def test_command_line_under_torchrun_gloo_world2(tmp_path, shard):
    """Two ranks (gloo, CPU).  rows: every chunk's rows are sharded over the ranks, rank 0 writes the
    files.  chrom: every rank runs whole chromosomes of its own (parse, classify, aggregate, genotype)
    into the shared directory and rank 0 merges.  Either way the merged VCF equals the single-process
    run."""
    import torch.multiprocessing as mp
    seg_dir, _, _, _ = _prepare(tmp_path)
    outs = {}
    for name in ("single", "sharded"):
        out = tmp_path / name
        (out / "segments").mkdir(parents=True)
        for f in os.listdir(seg_dir):
            (out / "segments" / f).write_bytes(open(os.path.join(seg_dir, f), "rb").read())
        outs[name] = out
    argv = lambda out: ["-o", str(out), "-b", "x.bam", "-m", "x.ckpt", "-g", str(tmp_path / "genome.fa"),   # noqa: E731
                        "-n", "S", "-s", "2", "--qname", "--shard", shard, "--debug"]
    assert step2.main(argv(outs["single"]), classifier=HashClassifier(),
                      genotype_for=lambda chrom: (lambda *a: ("0/1", 3, 4))) == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + 333
    procs = [ctx.Process(target=_step2_worker, args=(r, 2, port, argv(outs["sharded"]), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, 0), (1, 0)]
    a = (outs["single"] / "S.svision.s2.vcf").read_text()
    b = (outs["sharded"] / "S.svision.s2.vcf").read_text()
    assert a == b and a.count("\n") > 100
    if shard == "chrom":                                 # the ranks wrote disjoint chromosomes into ONE directory
        files = sorted(f for f in os.listdir(outs["sharded"] / "predict_results") if f.endswith(".vcf"))
        assert files == sorted(f for f in os.listdir(outs["single"] / "predict_results") if f.endswith(".vcf"))
        parts = step2.assign_chromosomes(["chr1", "chr2", "chrX"], str(outs["sharded"] / "segments"), 2)
        assert sorted(c for p in parts for c in p) == ["chr1", "chr2", "chrX"] and all(parts)

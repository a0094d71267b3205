"""ORACLE tooling: golden vectors for segment-pair generation.

In the build container, with the unmodified reference imported from /root/reference (pysam replaced by
``oracle/pysam_stub``):

1. runs the reference's collection stage on its demo BAM (as ``oracle/make_demo_rows.py`` does),
   captures the clusters handed to ``writer_cluster_to_file`` (src/collection/output_clusters.py:31)
   *before* ``get_segs_cords`` rewrites their alignments in place, and stores them in
   ``tests/golden/demo_clusters.json``; the text the reference then writes is
   ``tests/golden/demo_chr9.segments.bed`` (already committed) -- regenerated and compared here;
2. builds seeded synthetic signatures (1-7 alignments, reverse inner segments, zero-span and
   single-alignment cases), runs the reference's ``proc_one_cluster`` on them and stores inputs +
   the BED text in ``tests/golden/pairs_fuzz_golden.npz``.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_pairs_golden.py
"""
from __future__ import annotations

import copy
import json
import logging
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SVISION_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "pysam_stub"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

FUZZ_SEED, FUZZ_CLUSTERS = 31, 400


def cluster_to_dict(cl):
    return {"contig": cl.contig, "cstart": cl.cstart, "cend": cl.cend, "coverage": cl.coverage, "read_num": cl.read_num,
            "signatures": [{"qname": s.qname, "type": s.type, "mechanism": s.mechanism,
                            "bkps": [[int(v) for v in b[:3]] for b in s.bkps],
                            "aligns": [[int(a["ref_start"]), int(a["ref_end"]), int(a["q_start"]), int(a["q_end"]),
                                        bool(a["is_reverse"])] for a in s.sorted_aligns]} for s in cl.get_signatures()]}


def dict_to_cluster(d, Signature=None):
    """Rebuild a cluster-like object (for the reference when ``Signature`` is its class, else a plain namespace)."""
    sigs = []
    for s in d["signatures"]:
        aligns = [{"ref_start": a[0], "ref_end": a[1], "q_start": a[2], "q_end": a[3], "is_reverse": a[4]} for a in s["aligns"]]
        if Signature is not None:
            sigs.append(Signature(d["contig"], 0, 1, s["type"], s["qname"], aligns, copy.deepcopy(s["bkps"]), s["mechanism"]))
        else:
            sigs.append(types.SimpleNamespace(qname=s["qname"], type=s["type"], mechanism=s["mechanism"],
                                              bkps=s["bkps"], sorted_aligns=aligns))
    return types.SimpleNamespace(contig=d["contig"], cstart=d["cstart"], cend=d["cend"], coverage=d["coverage"],
                                 read_num=d["read_num"], get_signatures=lambda: sigs)


def synthetic_clusters(seed=FUZZ_SEED, n=FUZZ_CLUSTERS):
    rng = np.random.default_rng(seed)
    out = []
    pos = 50_000
    for c in range(n):
        width = int(10 ** rng.uniform(1.5, 4.5))
        sigs = []
        for r in range(int(rng.integers(1, 9))):
            n_aln = int(rng.choice([1, 2, 2, 2, 3, 3, 4, 5, 7]))
            ref, read = pos - int(rng.integers(0, 5000)), int(rng.integers(0, 3000))
            aligns = []
            for k in range(n_aln):
                ln = int(rng.integers(1, 6000))
                rev = bool(k not in (0, n_aln - 1) and rng.random() < 0.45) or bool(rng.random() < 0.05)
                aligns.append([ref, ref + ln - int(rng.integers(0, 2)), read, read + ln + int(rng.integers(-3, 4)), rev])
                kind = rng.random()
                if kind < 0.35:                       # deletion-like gap on the reference
                    ref += ln + int(rng.integers(0, width + 1))
                    read += ln + int(rng.integers(0, 3))
                elif kind < 0.7:                      # insertion-like gap on the read
                    ref += ln + int(rng.integers(0, 3))
                    read += ln + int(rng.integers(0, width + 1))
                elif kind < 0.85:                     # duplication-like jump back on the reference
                    ref += ln - int(rng.integers(0, width + 1))
                    read += ln
                else:                                 # co-linear continuation
                    d = int(rng.integers(0, 200))
                    ref += ln + d
                    read += ln + d
            if rng.random() < 0.03:                   # degenerate: every segment on one reference base
                aligns = [[pos, pos, a[2], a[3], a[4]] for a in aligns]
            bk = [[pos + int(rng.integers(-20, 20)), pos + width + int(rng.integers(-20, 20)), int(rng.integers(30, 9000))]
                  for _ in range(max(1, n_aln - 1))]
            sigs.append({"qname": f"r{c}/{r}", "type": "sigUncovered" if rng.random() < 0.1 else "sigGap",
                         "mechanism": ["None", "NHEJ+0", "FoSTeS+2"][int(rng.integers(0, 3))], "bkps": bk, "aligns": aligns})
        out.append({"contig": "chr2", "cstart": pos + float(rng.random()), "cend": pos + width + float(rng.random()),
                    "coverage": int(rng.integers(0, 80)), "read_num": len(sigs), "signatures": sigs})
        pos += width + int(rng.integers(1000, 30000))
    return out


def reference_lines(cluster_dicts, min_support, max_sv_size):
    """BED text written by the reference's proc_one_cluster under the writer's filters."""
    from src.collection.output_clusters import proc_one_cluster
    from src.collection.classes import Signature
    opt = types.SimpleNamespace(min_support=min_support, max_sv_size=max_sv_size, graph=False)
    lines = []
    for d in cluster_dicts:
        cl = dict_to_cluster(d, Signature)
        if int(cl.cend) - int(cl.cstart) > opt.max_sv_size or cl.read_num < opt.min_support:   # output_clusters.py:49-53
            continue
        lines.extend(proc_one_cluster(cl, opt)[1])
    return "".join(lines)


def main():
    logging.basicConfig(level=logging.WARNING)
    import src.collection.run_collection as RC
    captured = []
    original = RC.writer_cluster_to_file

    def capture(clusters, chrom, part_num, options):
        captured.extend(cluster_to_dict(c) for c in clusters)
        return original(clusters, chrom, part_num, options)

    RC.writer_cluster_to_file = capture
    bam = os.path.join(REF, "supports", "HG00733.svision.demo.bam")
    tmp = tempfile.mkdtemp(prefix="svx_pairs_")
    os.makedirs(os.path.join(tmp, "segments"), exist_ok=True)
    genome = os.path.join(tmp, "fake.fa")
    open(genome, "w").write(">chr9\nN\n")
    open(genome + ".fai", "w").write("chr9\t138394717\t6\t60\t61\n")
    opt = types.SimpleNamespace(
        genome=genome, out_path=tmp, sample="demo", min_support=5, min_mapq=10, min_sv_size=50,
        max_sv_size=1000000, patition_max_distance=5000, cluster_max_distance=0.3, hash=False,
        graph=False, contig=False, k_size=10, min_accept=50, max_hash_len=1000, qname=False,
        window_size=10000000, thread_num=1, debug=True)
    err = RC.run_detect(opt, bam, "chr9", 0, 70_000_000, 80_000_000)
    if err:
        raise RuntimeError(err)
    text = open(os.path.join(tmp, "segments", "chr9.segments.0.bed")).read()
    committed = open(os.path.join(ROOT, "tests", "golden", "demo_chr9.segments.bed")).read()
    assert text == committed, "the demo BED changed"
    assert reference_lines(captured, 5, 1000000) == committed        # the captured clusters reproduce it
    with open(os.path.join(ROOT, "tests", "golden", "demo_clusters.json"), "w") as f:
        json.dump({"min_support": 5, "max_sv_size": 1000000, "clusters": captured}, f, separators=(",", ":"))
    print("demo:", len(captured), "clusters,", sum(len(c["signatures"]) for c in captured), "signatures,",
          committed.count("\n"), "rows")

    fuzz = synthetic_clusters()
    fuzz_text = reference_lines(fuzz, 2, 20000)
    dst = os.path.join(ROOT, "tests", "golden", "pairs_fuzz_golden.npz")
    np.savez_compressed(dst, clusters=np.array(json.dumps(fuzz, separators=(",", ":"))), text=np.array(fuzz_text),
                        meta=np.array([FUZZ_SEED, FUZZ_CLUSTERS, 2, 20000]))
    print("fuzz:", len(fuzz), "clusters,", sum(len(c["signatures"]) for c in fuzz), "signatures,", fuzz_text.count("\n"),
          "rows ->", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()

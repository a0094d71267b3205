"""BASELINE config 4 / 5 as a whole-stream run on one GPU:

    python tools/stream_run.py [--rows 500000] [--profile hifi|ont] [--out DIR]

synthetic region-grouped stream -> 23-column BED text -> native parser (svx_bed_parse) -> encode +
classify on the GPU through the host entry (svx_classify) -> region aggregation, VCF records and
one-pass genotyping (svision_b200.calls) -> <out>/chr1.predict.s3.{vcf,score.txt}.  Prints one JSON
line with the rows/s of every stage.  Weights are synthetic (no checkpoint exists here), so the calls
are not biology; the run exercises formats and stage throughputs at whole-genome row counts."""
import argparse
import json
import os
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svision_b200 import bed, calls, classifier as C, sites, step2, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=500_000)
    ap.add_argument("--profile", default="hifi", choices=["hifi", "ont"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--chroms", type=int, default=0,
                    help="also run the whole Step 2 (svision_b200.step2: predict every chromosome on one "
                         "classifier, score range, merged VCF) over this many extra chromosomes of rows/10 rows")
    a = ap.parse_args()
    out_dir = a.out or tempfile.mkdtemp(prefix="svx_stream_")
    os.makedirs(out_dir, exist_ok=True)
    seed = sites.SEED_CONFIG4 if a.profile == "hifi" else sites.SEED_CONFIG5
    res = {"rows": a.rows, "profile": a.profile}

    t = time.perf_counter()
    table = sites.make_region_table(a.rows, seed=seed, profile=a.profile)
    aln = sites.make_alignments(table, seed=2)
    bed_path = os.path.join(out_dir, "chr1.segments.all.bed")
    with open(bed_path, "w") as f:
        f.write("\n".join(sites.table_to_bed_lines(table)) + "\n")
    res["generate_s"] = round(time.perf_counter() - t, 2)
    res["regions"] = len(set(table.region.tolist()))
    res["bed_mb"] = round(os.path.getsize(bed_path) / 1e6, 1)

    t = time.perf_counter()
    parsed = bed.read_segments_bed(bed_path)
    dt = time.perf_counter() - t
    res["bed_parse_rows_per_s"] = round(a.rows / dt)
    assert np.array_equal(parsed.rows, table.rows)

    clf = C.Classifier(weights.synthetic_weights(), device=0, max_batch=2048)
    clf.classify(parsed.rows[:4096])                                     # warm-up
    t = time.perf_counter()
    labels, probs = clf.classify(parsed.rows)
    dt = time.perf_counter() - t
    res["classify_sites_per_s"] = round(a.rows / dt)
    res["classify_s"] = round(dt, 3)
    res["label_histogram"] = np.bincount(labels, minlength=5).tolist()

    opt = types.SimpleNamespace(min_support=3, qname=False, min_sv_size=50, min_mapq=10, min_gt_depth=4,
                                homo_thresh=0.8, hete_thresh=0.2, bam_path="synthetic.bam")
    t = time.perf_counter()
    at = calls.AlignmentTable(aln["contig_length"], aln["reference_start"], aln["reference_end"],
                              aln["mapping_quality"], aln["is_unmapped"], aln["is_secondary"], aln["query_name"])
    res["alignment_index_s"] = round(time.perf_counter() - t, 3)
    t = time.perf_counter()
    records = calls.call_chromosome(parsed, labels, probs, opt, at)
    dt = time.perf_counter() - t
    res["calls_rows_per_s"] = round(a.rows / dt)
    res["calls_s"] = round(dt, 3)
    t = time.perf_counter()
    calls.write_chromosome(os.path.join(out_dir, "chr1.predict.s3"), records)
    res["write_s"] = round(time.perf_counter() - t, 3)
    res["records"] = len(records)
    total = a.rows / res["bed_parse_rows_per_s"] + res["classify_s"] + res["alignment_index_s"] + res["calls_s"] + res["write_s"]
    res["pipeline_rows_per_s"] = round(a.rows / total)
    # the same with GPU classification of chunk k+1 overlapping the record assembly of chunk k
    # (what predict.run_predict does): identical records
    t = time.perf_counter()
    streamed = calls.call_chromosome_streamed(parsed, clf.classify, opt, at)
    dt = time.perf_counter() - t
    assert [l for _, l in streamed] == [l for _, l in records]
    res["streamed_classify_plus_calls_s"] = round(dt, 3)
    total_s = a.rows / res["bed_parse_rows_per_s"] + dt + res["alignment_index_s"] + res["write_s"]
    res["pipeline_streamed_rows_per_s"] = round(a.rows / total_s)
    if a.chroms > 0:
        seg_dir, pred_dir = os.path.join(out_dir, "segments"), os.path.join(out_dir, "predict_results")
        os.makedirs(seg_dir, exist_ok=True)
        names, aligns, n_rows = [f"chr{k + 1}" for k in range(a.chroms)], {}, 0
        for k, chrom in enumerate(names):
            t_k = sites.make_region_table(max(a.rows // 10, 1000), seed=seed + 100 + k, profile=a.profile, contig=chrom)
            al = sites.make_alignments(t_k, seed=7 + k)
            aligns[chrom] = calls.AlignmentTable(al["contig_length"], al["reference_start"], al["reference_end"],
                                                 al["mapping_quality"], al["is_unmapped"], al["is_secondary"],
                                                 al["query_name"])
            with open(os.path.join(seg_dir, chrom + ".segments.all.bed"), "w") as f:
                f.write("\n".join(sites.table_to_bed_lines(t_k)) + "\n")
            n_rows += len(t_k)
        opt.sample, opt.out_path, opt.graph, opt.model_path = "synthetic", out_dir, False, "unused"
        t = time.perf_counter()
        merged = step2.run_step2(names, seg_dir, pred_dir, opt, classifier=clf, genotype_for=aligns.get,
                                 contigs=[(c, 250_000_000) for c in names])
        dt = time.perf_counter() - t
        res["step2"] = {"chromosomes": a.chroms, "rows": n_rows, "rows_per_s": round(n_rows / dt),
                        "merged_records": sum(1 for l in open(merged) if not l.startswith("#")),
                        "merged_vcf": os.path.basename(merged)}
    clf.close()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

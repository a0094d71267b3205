"""Host-side rows/s of the steps either side of the GPU path, CPU only.

    python tools/host_bench.py [--rows 100000] [--reference]

Prints one JSON line: BED parse (``bed.read_segments_bed``), post-classification calling
(``calls.call_chromosome`` with the one-pass ``AlignmentTable`` genotyper) and, with ``--reference`` in
the build container, the reference's own functions on the same stream (``oracle/make_calls_golden.py``
harness; its genotyper runs over an in-memory fake BAM, so its real per-record BAM re-open cost --
``genotype.py:22`` -- is NOT included: the ratio printed here is a lower bound)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svision_b200 import bed, calls, sites          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--reference", action="store_true")
    a = ap.parse_args()
    from oracle import make_calls_golden as G       # label synthesis only (tools/ is not the product path)
    table = sites.make_region_table(a.rows, seed=sites.SEED_CONFIG4)
    labels, probs = G.synthetic_labels(table, 1)
    aln = sites.make_alignments(table, seed=2)
    opt = G.options(3, False)
    out = {"rows": a.rows, "regions": len(set(table.region.tolist())), "alignments": int(aln["reference_start"].size)}

    with tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False) as f:
        f.write("\n".join(sites.table_to_bed_lines(table)) + "\n")
    t = time.perf_counter()
    parsed = bed.read_segments_bed(f.name)
    out["bed_parse_rows_per_s"] = round(a.rows / (time.perf_counter() - t))
    os.unlink(f.name)
    assert np.array_equal(parsed.rows, table.rows)

    t = time.perf_counter()
    at = calls.AlignmentTable(aln["contig_length"], aln["reference_start"], aln["reference_end"],
                              aln["mapping_quality"], aln["is_unmapped"], aln["is_secondary"], aln["query_name"])
    t_index = time.perf_counter() - t
    t = time.perf_counter()
    recs = calls.call_chromosome(table, labels, probs, opt, at)
    dt = time.perf_counter() - t
    out.update(records=len(recs), alignment_index_s=round(t_index, 3), calls_python_rows_per_s=round(a.rows / dt),
               calls_python_s=round(dt, 3))
    # the production route: table parsed from BED text -> svx_calls_aggregate (csrc/host_calls.cpp)
    t = time.perf_counter()
    recs_native = calls.call_chromosome(parsed, labels, probs, opt, at)
    dn = time.perf_counter() - t
    assert [l for _, l in recs_native] == [l for _, l in recs], "native aggregation differs from the Python route"
    out.update(calls_rows_per_s=round(a.rows / dn), calls_s=round(dn, 3))
    if a.reference:
        t = time.perf_counter()
        vcf, _, opens = G.reference_text(table, labels, probs, aln, opt)
        dr = time.perf_counter() - t
        assert vcf == "".join(l + "\n" for _, l in recs), "text differs from the reference's"
        out.update(reference_rows_per_s=round(a.rows / dr), reference_s=round(dr, 3), reference_bam_opens=opens,
                   speedup_lower_bound=round(dr / dt, 2))
    # §8(f) #4: signatures -> packed rows
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_pairs_golden", os.path.join(ROOT, "oracle", "make_pairs_golden.py"))
    PG = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(PG)
    from svision_b200 import pairs
    clusters = PG.synthetic_clusters(3, max(200, a.rows // 9))
    sig = pairs.SignatureTable.from_clusters([PG.dict_to_cluster(d) for d in clusters], 2, 20000)
    t = time.perf_counter()
    tb = pairs.generate_pairs(sig)
    dp = time.perf_counter() - t
    out.update(pair_signatures=len(sig), pair_rows=len(tb), pairs_rows_per_s=round(len(tb) / dp))
    if a.reference:
        t = time.perf_counter()
        ref_text = PG.reference_lines(clusters, 2, 20000)
        drp = time.perf_counter() - t
        assert ref_text == "".join(l + "\n" for l in pairs.to_bed_lines(tb))
        out.update(reference_pairs_rows_per_s=round(len(tb) / drp))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Whole Step 2 through the command-line entry, on 1 or N GPUs:

    python tools/step2_demo.py [--rows 200000] [--chroms 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
        tools/step2_demo.py --rows 200000 --chroms 3

Rank 0 prepares <out>/segments/<chrom>.segments.all.bed (synthetic region-grouped streams), a genome
.fai and a TF-format checkpoint of the synthetic weights (written with svision_b200.tf_bundle, read back
through the -m loader, so the TF-free reader is on the path), then every rank runs
``svision_b200.step2.main`` with the reference's flags (--shard auto: whole chromosomes per rank when
there are at least as many chromosomes as ranks, else the rows of every chunk over the ranks).
Genotypes come from synthetic alignment tables built by the rank that owns the chromosome (reading a
real BAM needs pysam, as in the reference).  Prints one JSON line on rank 0: rows/s of the whole step
and the digest of the merged VCF (equal for every world size and sharding mode)."""
import argparse
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from svision_b200 import calls, sites, step2, tf_bundle, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=200_000, help="rows per chromosome")
    ap.add_argument("--chroms", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--shard", default="auto", choices=["auto", "chrom", "rows"])
    ap.add_argument("--profile", default="hifi", choices=["hifi", "ont"],
                    help="segment-length profile of the synthetic rows (SURVEY 8(d): ont = BASELINE configs[4])")
    ap.add_argument("--contig", action="store_true", help="pass the reference's --contig flag (min_support 1)")
    ap.add_argument("--devices", default=None,
                    help="without torchrun: 'all' or '0,1,...' -- this ONE process drives these GPUs through svx_multi_*")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    out = a.out or os.path.join(tempfile.gettempdir(), "svx_step2_demo")
    names = [f"chr{k + 1}" for k in range(a.chroms)]
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    def table_of(k):
        seed = (sites.SEED_CONFIG4 if a.profile == "hifi" else sites.SEED_CONFIG5) + 10 * k
        return sites.make_region_table(a.rows, seed=seed, profile=a.profile, contig=names[k])

    # every rank prepares the segment files of chromosomes k = rank (mod world): Step 1's output
    tables = {}
    os.makedirs(os.path.join(out, "segments"), exist_ok=True)
    for k, chrom in enumerate(names):
        if k % world != rank:
            continue
        tables[chrom] = table_of(k)
        with open(os.path.join(out, "segments", chrom + ".segments.all.bed"), "w") as f:
            f.write("\n".join(sites.table_to_bed_lines(tables[chrom])) + "\n")
    if rank == 0:
        with open(os.path.join(out, "genome.fa.fai"), "w") as f:
            f.writelines(f"{c}\t250000000\t0\t70\t71\n" for c in names)
        tf_bundle.write_bundle(os.path.join(out, "model.ckpt"), weights.synthetic_weights())
    if world > 1:
        dist.barrier()
    # alignment tables (standing for the chromosome's BAM records) of the chromosomes this rank will
    # genotype: its own under chromosome sharding, all of them on rank 0 otherwise; built before the clock
    by_chrom = world > 1 and (a.shard == "chrom" or (a.shard == "auto" and a.chroms >= world))
    mine = step2.assign_chromosomes(names, os.path.join(out, "segments"), world)[rank] if by_chrom else \
        (names if rank == 0 else [])
    aligns = {}
    for chrom in mine:
        k = names.index(chrom)
        al = sites.make_alignments(tables[chrom] if chrom in tables else table_of(k), seed=7 + k)
        aligns[chrom] = calls.AlignmentTable(al["contig_length"], al["reference_start"], al["reference_end"],
                                             al["mapping_quality"], al["is_unmapped"], al["is_secondary"],
                                             al["query_name"])

    argv = ["-o", out, "-b", "synthetic.bam", "-m", os.path.join(out, "model.ckpt"), "-g", os.path.join(out, "genome.fa"),
            "-n", "demo", "-s", "3", "--debug", "--shard", a.shard] + (["--contig"] if a.contig else [])
    from svision_b200 import predict
    t = time.perf_counter()                                # -m loader (TF bundle, no TF) + weight repack + workspaces
    devices = None
    if a.devices and world == 1:
        import torch
        devices = list(range(torch.cuda.device_count())) if a.devices == "all" else [int(d) for d in a.devices.split(",")]
    clf = predict.get_classifier(os.path.join(out, "model.ckpt"), device=int(os.environ.get("LOCAL_RANK", rank)),
                                 devices=devices)
    # warm start: first launches, and the GPU out of its idle power state (a rank that had waited a few
    # seconds at the barrier measured 0.57 s for its first 65 k-row chunk against 0.19 s warm; a
    # whole-genome run pays that once)
    warm = sites.make_sites_p1(65536)
    clf.classify(warm)
    t_model = time.perf_counter() - t
    if world > 1:                                         # ranks finish their preparations at different times:
        dist.barrier()                                    # meet, warm again (an idle GPU has clocked down), meet
        clf.classify(warm)
        dist.barrier()
    t = time.perf_counter()
    rc = step2.main(argv, classifier=clf, genotype_for=aligns.get if (rank == 0 or by_chrom) else None)
    dt = time.perf_counter() - t
    if rank == 0:
        merged = os.path.join(out, "demo.svision.s1.vcf" if a.contig else "demo.svision.s3.vcf")
        text = open(merged, "rb").read()
        print(json.dumps({"world": world, "rc": rc, "shard": "chrom" if by_chrom else ("rows" if world > 1 else "none"),
                          "devices_in_this_process": len(devices) if devices else 1,
                          "profile": a.profile, "contig_mode": a.contig, "chromosomes": a.chroms, "rows": a.rows * a.chroms,
                          "model_load_s": round(t_model, 3), "step2_s": round(dt, 3), "rows_per_s": round(a.rows * a.chroms / dt),
                          "records": sum(1 for l in text.split(b"\n") if l and not l.startswith(b"#")),
                          "merged_sha256": hashlib.sha256(text).hexdigest()}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())

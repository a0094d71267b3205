"""Step 2 of the SVision driver ("CNN prediction", ``SVision:296-341``) around the B200 path:

    predict_chromosomes   <- predict_one_chrom + the multiprocessing pool      (SVision:298-323)
    score_range           <- cal_scores_max_min + np.max / np.min               (output.py:601-612, SVision:331-334)
    merge_chromosomes     <- merge_split_vcfs                                   (output.py:251-348)
    run_step2             <- the three in sequence                              (SVision:296-341)

What changes and why (SURVEY.md H6, §8(b) "Threading / processes"): the reference forks one worker
per chromosome, each building a TF session and restoring the checkpoint again, and never reads the
workers' error strings.  A CUDA context must not be forked and the model should be loaded once, so
the chromosomes run in one in-process loop over ONE classifier handle, and errors propagate.  What
does not change: the per-chromosome files ``<chrom>.predict.s<k>.{vcf,score.txt}`` and the merged
``<sample>.svision.s<k>.vcf`` are byte-identical to what the reference's functions write for the same
labels and scores (``tests/test_step2.py`` runs the reference's own ``merge_split_vcfs`` beside this).

The contig lines of the header come from the genome's ``.fai`` (name, length per line), which is what
``pysam.FastaFile`` serves at ``output.py:264-268``; no htslib is needed for that.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import predict as _predict

SVISION_VERSION = "1.4"                         # src/version.py (printed into the header: output.py:262)

# The fixed part of the header (output.py:270-299): (key, attributes) in file order.
_HEADER_FIELDS = (
    ("CHROM", 'CHROM=XXX,Description="Chromosome ID"'),
    ("POS", 'POS=XXX,Description="Start position of the SV described in this region"'),
    ("ID", 'ID=XXX,Description="ID of the SV described in this region"'),
    ("REF", 'REF=N,Description="Ref\'s sequence in that region, default=N"'),
    ("QUAL", 'QUAL=XXX,Description="The SV quality of the SV described in this region"'),
    ("ALT", 'ID=SV,Description="Simple SVs"'),
    ("ALT", 'ID=CSV,Description="Complex or nested SVs"'),
    ("FILTER", 'ID=Covered,Description="Covered mean the SV is spanned by reads"'),
    ("FILTER", 'ID=Uncovered,Description="UnCovered mean the SV is not spanned by reads"'),
    ("FILTER", 'ID=Clustered,Description="Clustered mean the SV is not spanned by reads, but can be cluster '
               'together with others"'),
    ("INFO", 'ID=END,Number=1,Type=Integer,Description="End position of the SV described in this region"'),
    ("INFO", 'ID=SVLEN,Number=1,Type=Integer,Description="Difference in length between REF and ALT alleles"'),
    ("INFO", 'ID=BKPS,Number=.,Type=String,Description="All breakpoints (length-start-end) in this region, '
             'where CSV might contain multiple breakpoints."'),
    ("INFO", 'ID=SVTYPE,Number=1,Type=String,Description="CNN predicted SV type, containing INS, DEL, DUP, tDUP '
             '(tandem duplication) and INV"'),
    ("INFO", 'ID=SUPPORT,Number=1,Type=Integer,Description="SV support number in this region"'),
    ("INFO", 'ID=READS,Number=.,Type=String,Description="SV support read names in this region"'),
)
_HEADER_GRAPH = (                                # only with --graph (output.py:288-292)
    ("INFO", 'ID=GraphID,Number=1,Type=String,Description="The corresponding graph id of isomorphic CSV graph '
             'structures"'),
    ("INFO", 'ID=GFA_FILE_PREFIX,Number=1,Type=String,Description="File name of CSV corresponding GFA file"'),
    ("INFO", 'ID=GFA_S,Number=1,Type=String,Description="Nodes contained in a CSV graph represented based on GFA '
             'format"'),
    ("INFO", 'ID=GFA_L,Number=1,Type=String,Description="Links contained in a CSV graph represented based on GFA '
             'format"'),
)
_HEADER_FORMAT = (
    ("FORMAT", 'ID=GT,Number=1,Type=String,Description="Genotype"'),
    ("FORMAT", 'ID=DR,Number=1,Type=Integer,Description="high-quality reference reads"'),
    ("FORMAT", 'ID=DV,Number=1,Type=Integer,Description="high-quality variant reads"'),
)


def contigs_from_fai(genome_path: str) -> List[Tuple[str, int]]:
    """``[(name, length), ...]`` in index order from ``<genome>.fai`` (what ``pysam.FastaFile(genome)
    .references`` / ``.get_reference_length`` return: output.py:264-268)."""
    out = []
    with open(genome_path + ".fai") as f:
        for line in f:
            cols = line.rstrip("\n").split("\t")
            if len(cols) >= 2 and cols[0]:
                out.append((cols[0], int(cols[1])))
    return out


def header_lines(contigs: Iterable[Tuple[str, int]], sample: str, graph: bool = False) -> List[str]:
    lines = ["##fileformat=VCFv4.3", f"##source=SVision v{SVISION_VERSION}"]
    lines += [f"##contig=<ID={name},length={length}>" for name, length in contigs]
    fields = _HEADER_FIELDS + (_HEADER_GRAPH if graph else ()) + _HEADER_FORMAT
    lines += [f"##{key}=<{attrs}>" for key, attrs in fields]
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + str(sample))
    return lines


def predict_paths(predict_dir: str, chrom: str, min_support) -> str:
    """Prefix of a chromosome's result files (SVision:301)."""
    return os.path.join(predict_dir, f"{chrom}.predict.s{min_support}")


def predict_chromosomes(chroms: Sequence[str], segments_dir: str, predict_dir: str, options,
                        classifier=None, genotype_for: Optional[Callable] = None) -> dict:
    """Runs ``Predict.run`` for every chromosome in order, in this process, on one classifier
    (``SVision:298-323`` used a fork pool of ``thread_num // 3`` workers).  ``genotype_for(chrom)``
    returns what :func:`svision_b200.predict.run_predict` accepts as ``genotype`` (default: the BAM of
    ``options.bam_path`` is read once per chromosome).  A chromosome without a segments file is
    skipped, as the reference's worker fails on it silently.  Returns ``{chrom: n_records}``."""
    os.makedirs(predict_dir, exist_ok=True)
    clf = classifier if classifier is not None else _predict.get_classifier(options.model_path)
    done = {}
    for chrom in chroms:
        bed_path = os.path.join(segments_dir, chrom + ".segments.all.bed")
        if not os.path.exists(bed_path):
            logging.warning("no segments file for %s (%s): skipped", chrom, bed_path)
            continue
        genotype = genotype_for(chrom) if genotype_for is not None else None
        done[chrom] = _predict.run_predict(bed_path, predict_paths(predict_dir, chrom, options.min_support),
                                           options, classifier=clf, chrom=chrom, genotype=genotype)
    return done


def score_range(predict_dir: str) -> Tuple[np.float64, np.float64]:
    """(max, min) over every ``*score.txt`` line that is not ``0`` (output.py:601-612, SVision:331-334).
    The reference prints 'Empty output in the score file' and exits when there is none; here that is a
    ``ValueError`` for the caller to report."""
    scores = []
    for name in os.listdir(predict_dir):
        if "score.txt" not in name:
            continue
        with open(os.path.join(predict_dir, name)) as f:
            for line in f:
                t = line.strip()
                if t == "0":
                    continue
                scores.append(float(t))
    if not scores:
        raise ValueError("no scores under " + predict_dir + ": nothing was called")
    return np.max(scores), np.min(scores)


def chromosome_summary(predict_dir: str, chrom: str, min_support) -> Tuple[int, int, Optional[float], Optional[float]]:
    """(records, distinct (POS, END) runs, max score, min score) of one chromosome's result files: what
    the merge needs to know about a chromosome before it can number and rescale the ones after it."""
    prefix = predict_paths(predict_dir, chrom, min_support)
    n = runs = 0
    prev = None
    with open(prefix + ".vcf") as f:
        for record in f:
            cols = record.split("\t", 8)
            key = (cols[1], cols[7].split(";", 1)[0][4:])
            if key != prev:
                prev = key
                runs += 1
            n += 1
    hi = lo = None
    with open(prefix + ".score.txt") as f:
        for line in f:
            t = line.strip()
            if t == "0":
                continue
            v = float(t)
            hi = v if hi is None or v > hi else hi
            lo = v if lo is None or v < lo else lo
    return n, runs, hi, lo


def merge_chromosome_records(predict_dir: str, chrom: str, min_support, serial: int, max_score, min_score,
                             out) -> Tuple[int, int]:
    """One chromosome's records with final IDs and rescaled QUAL written to ``out`` (output.py:307-345).
    ``serial`` is the ID of the last record written before this chromosome (-1 at the start); returns
    (the ID of this chromosome's last run, records written).  IDs count records whose (POS, END) differ from the previous
    record's; a record repeating the previous (POS, END) becomes ``<id>_<k>`` (output.py:318-331); the
    comparison starts afresh in every chromosome.  QUAL becomes
    ``int(100 - round((q - min) / (max - min), 2) * 100)``, or 100 when all scores are equal
    (output.py:334-341)."""
    span = max_score - min_score
    prev_key, sub, n = None, 1, 0
    with open(predict_paths(predict_dir, chrom, min_support) + ".vcf") as f:
        for record in f:
            cols = record.split("\t", 8)
            key = (cols[1], cols[7].split(";", 1)[0][4:])              # POS, END=<..>
            if key == prev_key:
                cols[2] = f"{serial}_{sub}"
                sub += 1
            else:
                prev_key, sub = key, 1
                serial += 1
                cols[2] = str(serial)
            q = 100
            if max_score != min_score:
                q = int(100 - (round((float(cols[5]) - min_score) / span, 2) * 100))
            cols[5] = str(q)
            out.write("\t".join(cols))
            n += 1
    return serial, n


def merge_chromosomes(predict_dir: str, merged_vcf_path: str, max_score, min_score, chroms: Sequence[str],
                      options, contigs: Optional[Iterable[Tuple[str, int]]] = None) -> int:
    """Header + every chromosome's records with final IDs and rescaled QUAL (output.py:251-348).
    Returns the number of records written."""
    if contigs is None:
        contigs = contigs_from_fai(options.genome)
    with open(merged_vcf_path, "w") as out:
        out.write("\n".join(header_lines(contigs, options.sample, bool(getattr(options, "graph", False)))) + "\n")
        serial, n = -1, 0
        for chrom in chroms:
            serial, k = merge_chromosome_records(predict_dir, chrom, options.min_support, serial, max_score, min_score, out)
            n += k
    return n


def run_step2(chroms: Sequence[str], segments_dir: str, predict_dir: str, options, classifier=None,
              genotype_for: Optional[Callable] = None, contigs=None) -> str:
    """``SVision:296-341``: predict every chromosome, then score range and merge.  Returns the path of
    ``<out_path>/<sample>.svision.s<min_support>.vcf``."""
    done = predict_chromosomes(chroms, segments_dir, predict_dir, options, classifier, genotype_for)
    hi, lo = score_range(predict_dir)
    merged = os.path.join(options.out_path, f"{options.sample}.svision.s{options.min_support}.vcf")
    merge_chromosomes(predict_dir, merged, hi, lo, [c for c in chroms if c in done], options, contigs)
    return merged


# ------------------------------------------------------------------------------------------------
# command line: the reference's flags (SVision:27-106), Step 2 only
# ------------------------------------------------------------------------------------------------
def parse_arguments(argv=None):
    """Same flags, destinations and defaults as the reference driver (``SVision:27-106``), so one
    command line serves both; flags that only steer Step 1 (collection) are accepted and ignored."""
    import argparse
    p = argparse.ArgumentParser(
        prog="python -m svision_b200.step2",
        description="SVision Step 2 (CNN prediction, genotyping, merged VCF) on the B200 path.  Expects "
                    "<out>/segments/<chrom>.segments.all.bed from Step 1 (run the reference with --debug to "
                    "keep them, or write them with svision_b200.pairs).")
    io = p.add_argument_group("Input/Output parameters")
    io.add_argument("-o", dest="out_path", type=os.path.abspath, required=True)
    io.add_argument("-b", dest="bam_path", type=os.path.abspath, required=True)
    io.add_argument("-m", dest="model_path", type=os.path.abspath, required=True)
    io.add_argument("-g", dest="genome", type=os.path.abspath, required=True)
    io.add_argument("-n", dest="sample", type=str, required=True)
    opt = p.add_argument_group("Optional parameters")
    opt.add_argument("-t", dest="thread_num", type=int, default=1)
    opt.add_argument("-s", dest="min_support", type=int, default=5)
    opt.add_argument("-c", dest="chrom", type=str, default=None)
    for flag in ("--hash", "--qname", "--graph", "--contig", "--debug"):
        opt.add_argument(flag, action="store_true", default=False)
    col = p.add_argument_group("Collect / cluster / hash parameters (Step 1; accepted for compatibility)")
    col.add_argument("--min_mapq", type=int, default=10)
    col.add_argument("--min_sv_size", type=int, default=50)
    col.add_argument("--max_sv_size", type=int, default=1000000)
    col.add_argument("--window_size", type=int, default=10000000)
    col.add_argument("--patition_max_distance", type=int, default=5000)
    col.add_argument("--cluster_max_distance", type=float, default=0.3)
    col.add_argument("--k_size", type=int, default=10)
    col.add_argument("--min_accept", type=int, default=50)
    col.add_argument("--max_hash_len", type=int, default=1000)
    pred = p.add_argument_group("Predict / genotype parameters")
    pred.add_argument("--batch_size", type=int, default=128, help="accepted and ignored: micro-batching is internal")
    pred.add_argument("--min_gt_depth", type=int, default=4)
    pred.add_argument("--homo_thresh", type=float, default=0.8)
    pred.add_argument("--hete_thresh", type=float, default=0.2)
    pred.add_argument("--device", type=int, default=0, help="CUDA device (not a reference flag)")
    pred.add_argument("--devices", type=str, default=None,
                      help="'all' or a comma-separated list: ONE process drives these GPUs through svx_multi_* "
                           "(the reference's single-process model; without torchrun; not a reference flag)")
    pred.add_argument("--shard", choices=("auto", "chrom", "rows"), default="auto",
                      help="under torchrun: whole chromosomes per rank (each rank parses, classifies, aggregates "
                           "and genotypes its own; rank 0 merges) or the rows of every chunk over the ranks; auto = "
                           "chromosomes when there are at least as many as ranks (not a reference flag)")
    options = p.parse_args(argv)
    if options.contig:                                   # SVision:161-162
        options.min_support = 1
    return options


def chromosomes_with_segments(segments_dir: str, contigs: Sequence[Tuple[str, int]], only: Optional[str] = None):
    """Chromosomes that have a segments file, in the genome's contig order (the reference walks the
    FASTA's contigs, ``SVision:167-234``); ``only`` = the ``-c`` flag (``chr1`` or ``chr1:a-b``)."""
    want = only.split(":")[0] if only else None
    return [name for name, _ in contigs
            if (want is None or name == want) and os.path.exists(os.path.join(segments_dir, name + ".segments.all.bed"))]


def assign_chromosomes(chroms: Sequence[str], segments_dir: str, world: int) -> List[List[str]]:
    """Whole chromosomes to ranks, longest segments file first onto the least-loaded rank (the
    reference's pool hands one chromosome to each free worker: SVision:311-323).  Deterministic, so
    every rank computes the same assignment; each rank's list keeps the genome's contig order."""
    size = {c: os.path.getsize(os.path.join(segments_dir, c + ".segments.all.bed")) for c in chroms}
    load = [0] * world
    owner = {}
    for c in sorted(chroms, key=lambda c: (-size[c], chroms.index(c))):
        r = min(range(world), key=lambda k: (load[k], k))
        owner[c] = r
        load[r] += size[c]
    return [[c for c in chroms if owner[c] == r] for r in range(world)]


def main(argv=None, classifier=None, genotype_for: Optional[Callable] = None) -> int:
    options = parse_arguments(argv)
    logging.basicConfig(level=logging.DEBUG if os.environ.get("SVX_STEP2_DEBUG") else logging.INFO,
                        format="%(asctime)s %(message)s")
    segments_dir = os.path.join(options.out_path, "segments")
    predict_dir = os.path.join(options.out_path, "predict_results")
    contigs = contigs_from_fai(options.genome)
    chroms = chromosomes_with_segments(segments_dir, contigs, options.chrom)
    if not chroms:
        logging.error("no <chrom>.segments.all.bed under %s", segments_dir)
        return 1
    # under torchrun (one process per GPU) either whole chromosomes go to the ranks (every rank parses,
    # classifies, aggregates and genotypes its own chromosomes into the shared predict_results directory;
    # rank 0 merges) or the rows of every chunk are sharded over the ranks (every rank walks the same
    # chromosomes so that the collectives line up, rank 0 owns the output files)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    by_chrom = world > 1 and (options.shard == "chrom" or (options.shard == "auto" and len(chroms) >= world))
    scratch = None
    if world > 1:
        import tempfile
        import torch
        import torch.distributed as dist
        from . import sharded
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl" if (classifier is None and torch.cuda.is_available() and not by_chrom)
                                    else "gloo")
        if classifier is None:
            options.device = int(os.environ.get("LOCAL_RANK", rank))
        if rank != 0 and not by_chrom:                    # same work, throw-away files, no BAM reads
            scratch = tempfile.mkdtemp(prefix=f"svx_step2_rank{rank}_")
            predict_dir = os.path.join(scratch, "predict_results")
            options.out_path = scratch
            if genotype_for is None:
                genotype_for = lambda chrom: (lambda *a: ("./.", 0, 0))    # noqa: E731
    if classifier is None:
        devices = None
        if options.devices and world == 1:
            import torch
            devices = list(range(torch.cuda.device_count())) if options.devices == "all" else \
                [int(d) for d in options.devices.split(",") if d.strip()]
        classifier = _predict.get_classifier(options.model_path, device=options.device, devices=devices)
    try:
        if by_chrom:
            import time
            t0 = time.perf_counter()
            mine = assign_chromosomes(chroms, segments_dir, world)[rank]
            done = predict_chromosomes(mine, segments_dir, predict_dir, options, classifier, genotype_for)
            t1 = time.perf_counter()
            # the merge in parallel too: IDs run on through the chromosomes and QUAL needs the global score
            # range, so the ranks first exchange (records, ID runs, max, min) per chromosome, then every rank
            # writes the final text of its own chromosomes and rank 0 only concatenates
            mine_done = [c for c in mine if c in done]
            summary = {c: chromosome_summary(predict_dir, c, options.min_support) for c in mine_done}
            everyone = [None] * world
            dist.all_gather_object(everyone, summary)                    # also the barrier: all files are written
            t2 = time.perf_counter()
            info = {c: v for part in everyone for c, v in part.items()}
            order = [c for c in chroms if c in info]
            his = [v[2] for v in info.values() if v[2] is not None]
            if not his:
                raise ValueError("no scores under " + predict_dir + ": nothing was called")
            hi, lo = np.max(his), np.min([v[3] for v in info.values() if v[3] is not None])
            serial = -1
            for c in order:
                if c in summary:
                    with open(predict_paths(predict_dir, c, options.min_support) + ".final.vcf", "w") as out:
                        merge_chromosome_records(predict_dir, c, options.min_support, serial, hi, lo, out)
                serial += info[c][1]
            dist.barrier()
            merged = os.path.join(options.out_path, f"{options.sample}.svision.s{options.min_support}.vcf")
            if rank == 0:
                with open(merged, "wb") as out:
                    out.write(("\n".join(header_lines(contigs, options.sample, bool(getattr(options, "graph", False)))) + "\n").encode())
                    for c in order:
                        part = predict_paths(predict_dir, c, options.min_support) + ".final.vcf"
                        with open(part, "rb") as f:
                            out.write(f.read())
                        os.remove(part)
            t3 = time.perf_counter()
            dist.barrier()
            logging.info("rank %d: own chromosomes %s in %.3f s, waited %.3f s for the others, merge %.3f s",
                         rank, ",".join(mine), t1 - t0, t2 - t1, t3 - t2)
        else:
            if world > 1:
                classifier = sharded.ShardedClassifier(classifier)
            merged = run_step2(chroms, segments_dir, predict_dir, options, classifier, genotype_for, contigs)
    except ValueError as e:                               # 'Empty output in the score file' (SVision:374-376)
        logging.error("%s", e)
        return 1
    finally:
        if scratch is not None:
            import shutil
            shutil.rmtree(scratch, ignore_errors=True)
    if rank == 0:
        logging.info("[Prediction finished] %d chromosome(s) on %d GPU(s) -> %s", len(chroms), world, merged)
        if not options.debug:                             # SVision:370-372 removes the intermediates
            import shutil
            shutil.rmtree(predict_dir, ignore_errors=True)
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())

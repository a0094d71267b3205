"""ORACLE tooling -- derive REAL-DATA encoder inputs by running the reference's own collection
stage (``src/collection/run_collection.run_detect`` -> ``analyze_alignments`` ->
``partition_and_cluster`` -> ``writer_cluster_to_file``) on its demo BAM
(``supports/HG00733.svision.demo.bam``) with the pure-Python pysam stand-in of
``oracle/pysam_stub`` (pysam/htslib are not installable in this image), then encode the rows with
the reference's ``BatchGenerator`` and store rows + lit-pixel codes in
``tests/golden/demo_rows_golden.npz``.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_demo_rows.py

The FASTA is a deterministic pseudo-sequence, so breakpoint left-shifting differs from a GRCh38
run: these are realistic *encoder inputs* (SURVEY.md §4 item 2), not a VCF golden."""
from __future__ import annotations

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
sys.path.insert(0, os.path.join(HERE, "pysam_stub"))      # `import pysam` -> the stand-in
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from svision_b200 import bed  # noqa: E402
from oracle import make_golden  # noqa: E402


def main():
    logging.basicConfig(level=logging.WARNING)
    from src.collection.run_collection import run_detect          # reference code, imported not copied
    bam = os.path.join(REF, "supports", "HG00733.svision.demo.bam")
    tmp = tempfile.mkdtemp(prefix="svx_demo_")
    os.makedirs(os.path.join(tmp, "segments"), exist_ok=True)
    genome = os.path.join(tmp, "fake.fa")
    open(genome, "w").write(">chr9\nN\n")
    open(genome + ".fai", "w").write("chr9\t138394717\t6\t60\t61\n")
    opt = types.SimpleNamespace(
        genome=genome, out_path=tmp, sample="demo", min_support=5, min_mapq=10, min_sv_size=50,
        max_sv_size=1000000, patition_max_distance=5000, cluster_max_distance=0.3, hash=False,
        graph=False, contig=False, k_size=10, min_accept=50, max_hash_len=1000, qname=False,
        window_size=10000000, thread_num=1, debug=True)
    rows_all = []
    for part, (s, e) in enumerate([(70_000_000, 80_000_000)]):
        err = run_detect(opt, bam, "chr9", part, s, e)
        if err:
            raise RuntimeError(err)
        path = os.path.join(tmp, "segments", f"chr9.segments.{part}.bed")
        if not os.path.exists(path):
            cands = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp) for f in fs if f.endswith(".bed")]
            raise RuntimeError(f"no BED written; found {cands}")
        import shutil
        shutil.copy(path, os.path.join(ROOT, "tests", "golden", "demo_chr9.segments.bed"))
        table = bed.read_segments_bed(path)
        print(f"part {part}: {len(table)} rows, {len(set(table.region))} regions")
        rows_all.append(table.rows)
    rows = np.concatenate(rows_all)
    imgs = make_golden.reference_images(rows)
    off, codes = make_golden.images_to_codes(imgs)
    dst = os.path.join(ROOT, "tests", "golden", "demo_rows_golden.npz")
    np.savez_compressed(dst, rows=rows, offsets=off, codes=codes,
                        meta=np.array(["source=reference run_detect on supports/HG00733.svision.demo.bam "
                                       "(chr9:70-80Mb, -s 5) with oracle/pysam_stub; images by reference "
                                       "BatchGenerator"]))
    print("wrote", dst, rows.shape)


if __name__ == "__main__":
    main()

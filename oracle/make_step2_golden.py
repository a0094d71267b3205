"""ORACLE tooling (build container only): pins the merged-VCF text of tests/test_step2.py.

Runs the REFERENCE's own ``cal_scores_max_min`` + ``merge_split_vcfs`` (src/network/output.py:251-348,
601-612, imported unmodified from /root/reference; only ``pysam.FastaFile`` is replaced by a reader of the
``.fai``) over the per-chromosome files of the test's synthetic three-chromosome run and writes
``tests/golden/step2_merged.sha256``: the SHA-256 of the whole merged file and of its records alone.
The GPU box has no reference tree; there the product's text is compared with these digests."""
import hashlib
import os
import pathlib
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import reference_loader as RL                      # noqa: E402
from svision_b200 import step2                                 # noqa: E402
import test_step2 as T                                         # noqa: E402


def main():
    with tempfile.TemporaryDirectory() as d:
        tmp = pathlib.Path(d)
        seg_dir, pred_dir, clf, tables = T._prepare(tmp)
        opt = T._options(tmp)
        chroms = ["chr1", "chr2", "chrX"]
        step2.predict_chromosomes(chroms, seg_dir, pred_dir, opt, classifier=clf, genotype_for=tables.get)

        class FastaFile:
            def __init__(self, path):
                self._c = step2.contigs_from_fai(path)
                self.references = [n for n, _ in self._c]

            def get_reference_length(self, name):
                return dict(self._c)[name]

        with RL.reference_modules() as ref:
            ref.output.pysam = types.SimpleNamespace(FastaFile=FastaFile)
            scores = ref.output.cal_scores_max_min(pred_dir)
            out = str(tmp / "reference_merged.vcf")
            ref.output.merge_split_vcfs(pred_dir, out, np.max(scores), np.min(scores), chroms, opt)
        text = open(out).read()
        body = "".join(l + "\n" for l in text.split("\n") if l and not l.startswith("#"))
        with open(os.path.join(ROOT, "tests", "golden", "step2_merged.sha256"), "w") as f:
            f.write(hashlib.sha256(text.encode()).hexdigest() + " " + hashlib.sha256(body.encode()).hexdigest() + "\n")
        print("records:", body.count("\n"))


if __name__ == "__main__":
    main()

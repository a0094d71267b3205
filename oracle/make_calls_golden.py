"""ORACLE tooling: golden VCF text for the post-classification step.

Runs, in the build container, the reference's own ``get_region_potential_svtypes`` +
``write_results_to_vcf`` + ``genotyper`` (imported unmodified from /root/reference through
``oracle/reference_loader.py``) over a seeded synthetic region stream with seeded class labels and
scores, and stores the text they print in ``tests/golden/calls_golden.npz``.  The per-row loop that
feeds them is ``svision_b200.predict.replay_rows`` (the restated ``predict.py:213-300``; the loop itself
cannot be imported because ``Predict.run`` needs TensorFlow).

    python oracle/make_calls_golden.py
"""
from __future__ import annotations

import io
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_loader as RL                      # noqa: E402
from svision_b200 import predict as P, sites                   # noqa: E402

N_ROWS, TABLE_SEED, LABEL_SEED, ALN_SEED = 6000, 424242, 99, 5


def synthetic_labels(table, seed):
    """Class labels with region-level structure (so candidates reach min_support) + float32 scores."""
    rng = np.random.default_rng(seed)
    n = len(table)
    labels = np.empty(n, dtype=np.int32)
    region_main, region_minor = {}, {}
    for i in range(n):
        reg = table.region[i]
        if reg not in region_main:
            region_main[reg] = int(rng.choice(5, p=[0.35, 0.35, 0.05, 0.1, 0.15]))
            region_minor[reg] = int(rng.choice(5, p=[0.05, 0.1, 0.25, 0.3, 0.3]))
        if "m" in table.read_num[i]:
            labels[i] = region_main[reg] if rng.random() < 0.85 else int(rng.integers(0, 5))
        else:
            labels[i] = region_minor[reg] if rng.random() < 0.75 else int(rng.integers(0, 5))
    logits = rng.normal(0, 1.5, size=(n, 5))
    logits[np.arange(n), labels] = logits.max(1) + 0.2 + 3.0 * rng.random(n)     # the label is the arg-max
    p = np.exp(logits - logits.max(1, keepdims=True))
    probs = (p / p.sum(1, keepdims=True)).astype(np.float32)
    assert (probs.argmax(1) == labels).all()
    return labels, probs


def options(min_support=3, qname=True, min_sv_size=50):
    return types.SimpleNamespace(min_support=min_support, qname=qname, min_sv_size=min_sv_size, min_mapq=10,
                                 min_gt_depth=4, homo_thresh=0.8, hete_thresh=0.2, bam_path="synthetic.bam")


def reference_text(table, labels, probs, alignments, opt):
    """(vcf text, score text) printed by the reference's functions."""
    with RL.reference_modules() as ref:
        fake = RL.FakePysam(alignments)
        ref.genotype.pysam = fake
        pred = ref.Predict("chr1", "unused")
        vcf, score = io.StringIO(), io.StringIO()

        def flush(region, reads_dict, names, sig_types, sig_scores, class_scores, mechs):
            ref.write_results_to_vcf(vcf, score, pred.get_region_potential_svtypes(reads_dict), region, names,
                                     sig_types, sig_scores, class_scores, mechs, opt)

        P.replay_rows(table, labels, probs, flush)
        return vcf.getvalue(), score.getvalue(), fake.opens


def main():
    table = sites.make_region_table(N_ROWS, seed=TABLE_SEED)
    labels, probs = synthetic_labels(table, LABEL_SEED)
    aln = sites.make_alignments(table, seed=ALN_SEED)
    out = {}
    for tag, opt in (("s3_qname", options(3, True)), ("s1", options(1, False)), ("s5_min200", options(5, False, 200))):
        vcf, score, opens = reference_text(table, labels, probs, aln, opt)
        out[f"{tag}_vcf"] = np.array(vcf)
        out[f"{tag}_score"] = np.array(score)
        print(tag, "records", vcf.count("\n"), "BAM opens by the reference genotyper", opens)
    out["labels"], out["probs"] = labels, probs
    out["meta"] = np.array([N_ROWS, TABLE_SEED, LABEL_SEED, ALN_SEED])
    path = os.path.join(ROOT, "tests", "golden", "calls_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

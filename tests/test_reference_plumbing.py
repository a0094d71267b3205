"""CPU, build container only: the drop-in claim end to end (BASELINE config 1 substitute).

``svision_b200.predict.run_predict`` is run on the real demo BED with the REFERENCE's own, unmodified
``Predict.get_region_potential_svtypes`` (src/network/predict.py:29-145), ``write_results_to_vcf``
(src/network/output.py:469-598) and ``genotyper`` (src/network/genotype.py:17-73) imported from
``/root/reference``; only the absent third-party imports are stubbed (tensorflow, bs4 -- unused on this
path -- and pysam via oracle/pysam_stub).  The classifier is the CPU oracle here (no GPU in the build
container); the GPU classifier is checked against that oracle on the same rows in tests/test_gpu_cnn.py.
Weights are synthetic, so the SV *types* are meaningless: this test is about plumbing and formats."""
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("SVISION_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "network")),
                                reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def reference_modules():
    sys.path.insert(0, ROOT)
    from oracle import reference_loader as RL
    with RL.reference_modules() as ref:
        yield ref.Predict, ref.write_results_to_vcf


def test_run_predict_with_reference_aggregation_and_vcf_writer(reference_modules, synthetic_weights, tmp_path):
    RefPredict, write_results_to_vcf = reference_modules
    from oracle import alexnet, encoder_c
    from svision_b200 import predict as P

    class OracleClassifier:                       # stands in for svision_b200.Classifier on CPU
        def classify(self, rows):
            l, p, _ = alexnet.classify(encoder_c.encode_f32(rows), synthetic_weights, torch.float32, batch=64)
            return l.astype(np.int32), p.astype(np.float32)

    bed_path = os.path.join(HERE, "golden", "demo_chr9.segments.bed")
    opt = types.SimpleNamespace(
        model_path="unused", batch_size=128, min_support=5, min_mapq=10, qname=False, graph=False,
        contig=False, min_gt_depth=4, homo_thresh=0.8, hete_thresh=0.2, min_sv_size=50,
        max_sv_size=1000000, sample="demo", out_path=str(tmp_path), genome="/nonexistent.fa",
        bam_path=os.path.join(REF, "supports", "HG00733.svision.demo.bam"))
    ref = RefPredict("chr9", bed_path)
    prefix = str(tmp_path / "chr9.predict.s5")
    torch.set_num_threads(4)
    n = P.run_predict(bed_path, prefix, opt, aggregate=ref.get_region_potential_svtypes,
                      write=write_results_to_vcf, classifier=OracleClassifier(), chrom="chr9")
    assert n == 11                                              # 11 clusters in the demo window
    vcf = [l.rstrip("\n").split("\t") for l in open(prefix + ".vcf")]
    scores = [float(l) for l in open(prefix + ".score.txt")]
    assert 1 <= len(vcf) <= 11 and len(scores) == len(vcf)      # regions with no call write nothing
    for rec, sc in zip(vcf, scores):
        assert len(rec) == 10 and rec[0] == "chr9" and rec[4] in ("<SV>", "<CSV>")
        assert float(rec[5]) == sc
        info = dict(kv.split("=", 1) for kv in rec[7].split(";"))
        assert {"END", "SVLEN", "SVTYPE", "SUPPORT", "BKPS"} <= set(info)
        assert int(info["END"]) >= int(rec[1]) and int(info["SUPPORT"]) >= 1
        assert set(info["SVTYPE"].split("+")) <= {"DEL", "INS", "INV", "DUP", "tDUP"}
        assert rec[8] == "GT:DR:DV" and len(rec[9].split(":")) == 3

    # default mode: this package's own aggregation / records / one-pass genotyper (svision_b200.calls),
    # BAM read once through the same pysam stand-in -> byte-identical files
    prefix2 = str(tmp_path / "own" / "chr9.predict.s5")
    os.makedirs(os.path.dirname(prefix2))
    m = P.run_predict(bed_path, prefix2, opt, classifier=OracleClassifier(), chrom="chr9")
    assert m == len(vcf)
    assert open(prefix2 + ".vcf").read() == open(prefix + ".vcf").read()
    assert open(prefix2 + ".score.txt").read() == open(prefix + ".score.txt").read()

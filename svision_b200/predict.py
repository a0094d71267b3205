"""Drop-in body for the reference's ``Predict.run`` (``src/network/predict.py:148-303``).

``run_predict`` parses the BED once (:mod:`svision_b200.bed`), classifies every row on the GPU in
one call (:class:`svision_b200.classifier.Classifier`; the model is loaded once per process, not
once per chromosome as ``predict.py:179-184`` does), then replays the reference's per-row
bookkeeping (``predict.py:213-300``) unchanged in meaning and hands each finished region to the
reference's own aggregation and VCF writer, which stay where they are:

    aggregate = Predict.get_region_potential_svtypes      (predict.py:29-145)
    write     = src.network.output.write_results_to_vcf    (output.py:469-598)

Both are *injected* so nothing of the reference is copied here; INTEGRATION.md shows the patch.
Output files are the reference's: ``<prefix>.vcf`` and ``<prefix>.score.txt`` (predict.py:157-158).
"""
from __future__ import annotations

import logging
from typing import Callable, Optional

import numpy as np

from . import bed as _bed

_CLASSIFIER_CACHE = {}


def get_classifier(model_path, device: int = 0, max_batch: int = 2048):
    """Process-wide classifier for ``-m model_path`` (amortises the checkpoint load)."""
    from .classifier import Classifier
    key = (str(model_path), int(device))
    if key not in _CLASSIFIER_CACHE:
        _CLASSIFIER_CACHE[key] = Classifier(model_path, device=device, max_batch=max_batch)
    return _CLASSIFIER_CACHE[key]


def replay_rows(table: "_bed.SegmentsTable", labels: np.ndarray, probs: np.ndarray,
                flush: Callable[..., None]) -> int:
    """The per-row loop of ``predict.py:213-300``.  ``flush(region, reads_dict, read_num_name_pair,
    sig_types, sig_score_pair, predict_scores, sig_mechanisms_pair)`` is called for every finished
    region and once at the end (as the reference does).  Returns the number of regions flushed."""
    assert probs.dtype == np.float32, "scores must stay numpy.float32 (SURVEY §8(b))"
    reads_dict, read_num_name_pair, sig_score_pair, sig_mechanisms_pair = {}, {}, {}, {}
    sig_types, predict_scores = [], []
    last_region = ""
    flushed = 0
    for i in range(len(table)):
        read_num = str(table.read_num[i])
        region = str(table.region[i])
        pred = int(labels[i])
        # v1.0.1 rule: a forward signature cannot be an inversion (predict.py:229-231)
        if str(table.forward[i]) == "True" and pred == 2:
            continue
        if region != last_region:                                     # predict.py:235-247
            if last_region != "":
                flush(last_region, reads_dict, read_num_name_pair, sig_types, sig_score_pair,
                      predict_scores, sig_mechanisms_pair)
                flushed += 1
            last_region = region
            reads_dict, read_num_name_pair, sig_score_pair, sig_mechanisms_pair = {}, {}, {}, {}
            sig_types, predict_scores = [], []
        key = read_num.replace("m", "")
        read_num_name_pair[key] = str(table.read_name[i])
        sig_types.append(str(table.sig_type[i]))
        predict_scores.append(round(probs[i][pred], 2))               # numpy.float32, predict.py:251
        sig_score_pair[key] = str(table.sig_score[i])
        sig_mechanisms_pair[key] = str(table.mechanism[i])
        bkp = [int(table.bkp_start[i]), int(table.bkp_end[i]), int(table.bkp_len[i])]
        if "m" not in read_num:                                       # predict.py:279-288
            if pred == 0 or pred == 1:          # only main segments may be called INS/DEL
                continue
            reads_dict.setdefault(read_num, {})[pred] = bkp
        else:                                                         # predict.py:290-296
            reads_dict.setdefault(key, {})[pred] = bkp
    flush(last_region, reads_dict, read_num_name_pair, sig_types, sig_score_pair, predict_scores,
          sig_mechanisms_pair)
    return flushed + 1


def run_predict(segments_out_file: str, out_path_prefix: str, options, aggregate: Callable,
                write: Callable, classifier=None, chrom: Optional[str] = None) -> int:
    """Body of ``Predict.run``.  ``options`` is the reference's argparse namespace (uses
    ``model_path``; ``batch_size`` is accepted and ignored: micro-batching is internal).
    Errors propagate (the reference swallows them: ``SVision:306-309``)."""
    table = _bed.read_segments_bed(segments_out_file)
    clf = classifier if classifier is not None else get_classifier(options.model_path)
    if chrom:
        logging.info("Predicting " + chrom)                           # predict.py:204
    labels, probs = clf.classify(table.rows)
    with open(out_path_prefix + ".score.txt", "w") as score_out, \
            open(out_path_prefix + ".vcf", "w") as vcf_out:

        def flush(region, reads_dict, read_num_name_pair, sig_types, sig_score_pair,
                  predict_scores, sig_mechanisms_pair):
            write(vcf_out, score_out, aggregate(reads_dict), region, read_num_name_pair, sig_types,
                  sig_score_pair, predict_scores, sig_mechanisms_pair, options)

        return replay_rows(table, labels, probs, flush)


class Predict:
    """Same constructor and ``run`` signature as the reference class (predict.py:14-27,148)."""

    def __init__(self, chrom, segments_out_file, aggregate: Callable = None, write: Callable = None,
                 classifier=None):
        self.segments_out_file = segments_out_file
        self.chrom = chrom
        self.num_classes = 5
        self._aggregate, self._write, self._classifier = aggregate, write, classifier

    def run(self, out_path_prefix, options):
        if self._aggregate is None or self._write is None:
            raise RuntimeError("Predict needs the reference's get_region_potential_svtypes and "
                               "write_results_to_vcf injected (see INTEGRATION.md)")
        return run_predict(self.segments_out_file, out_path_prefix, options, self._aggregate,
                           self._write, self._classifier, self.chrom)

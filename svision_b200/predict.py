"""Drop-in body for the reference's ``Predict.run`` (``src/network/predict.py:148-303``).

``run_predict`` parses the BED once (:mod:`svision_b200.bed`), classifies every row on the GPU in
one call (:class:`svision_b200.classifier.Classifier`; the model is loaded once per process, not
once per chromosome as ``predict.py:179-184`` does), then turns labels and scores into VCF records.
Two ways, same text:

* default: :mod:`svision_b200.calls` -- this package's own per-row loop, region aggregation, type
  refinement, record assembly and a genotyper that reads the BAM once per chromosome
  (byte-identical to the reference's text on the golden stream);
* injected: the reference's own functions stay where they are and are passed in,

      aggregate = Predict.get_region_potential_svtypes      (predict.py:29-145)
      write     = src.network.output.write_results_to_vcf    (output.py:469-598)

  with :func:`replay_rows` feeding them (INTEGRATION.md shows the patch).

Output files are the reference's: ``<prefix>.vcf`` and ``<prefix>.score.txt`` (predict.py:157-158).
"""
from __future__ import annotations

import logging
from typing import Callable, Optional

import numpy as np

from . import bed as _bed

_CLASSIFIER_CACHE = {}


def get_classifier(model_path, device: int = 0, max_batch: int = 8192, devices=None):
    """Process-wide classifier for ``-m model_path`` (amortises the checkpoint load).  ``devices`` with
    more than one entry gives ONE object driving all of them from this process (``MultiClassifier``,
    C-ABI ``svx_multi_*``): the reference's single-process model (``SVision:296-341``) on a multi-GPU box."""
    from .classifier import Classifier, MultiClassifier
    devs = tuple(int(d) for d in devices) if devices is not None else (int(device),)
    key = (str(model_path), devs)
    if key not in _CLASSIFIER_CACHE:
        if len(devs) > 1:
            _CLASSIFIER_CACHE[key] = MultiClassifier(model_path, devices=list(devs), max_batch=max_batch)
        else:
            _CLASSIFIER_CACHE[key] = Classifier(model_path, device=devs[0], max_batch=max_batch)
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
    complement = (np.asarray(table.flags) & _bed.FLAG_COMPLEMENT) != 0
    for i in range(len(table)):
        # a row whose label contains 'complement' anywhere is skipped outright (predict.py:214: the
        # reference's pad rows, but also any read / contig / mechanism named like that)
        if complement[i]:
            continue
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


def run_predict(segments_out_file: str, out_path_prefix: str, options, aggregate: Callable = None,
                write: Callable = None, classifier=None, chrom: Optional[str] = None, genotype=None) -> int:
    """Body of ``Predict.run``.  ``options`` is the reference's argparse namespace (uses
    ``model_path``, ``min_support``, ``min_sv_size``, ``qname``, ``bam_path``, ``min_mapq``,
    ``min_gt_depth``, ``homo_thresh``, ``hete_thresh``; ``batch_size`` is accepted and ignored:
    micro-batching is internal).  Errors propagate (the reference swallows them: ``SVision:306-309``).

    Without ``aggregate``/``write`` the records come from :mod:`svision_b200.calls`; ``genotype`` is
    then a :class:`calls.AlignmentTable`, a callable ``(candidate, read_names, options) -> (GT, DR, DV)``,
    or None to load ``options.bam_path`` once for ``chrom`` (needs pysam, as the reference does).
    Returns the number of regions flushed (injected mode) or of records written (default mode)."""
    import time
    t0 = time.perf_counter()
    table = _bed.read_segments_bed(segments_out_file)
    t_parse = time.perf_counter() - t0
    clf = classifier if classifier is not None else get_classifier(options.model_path)
    if chrom:
        logging.info("Predicting " + chrom)                           # predict.py:204
    if aggregate is None and write is None:
        from . import calls
        if genotype is None:
            contig = chrom if chrom else (str(table.region[0]).split("+")[0] if len(table) else None)
            genotype = calls.AlignmentTable.from_bam(options.bam_path, contig) if contig else (lambda *a: ("./.", 0, 0))
        # classification of the next chunk of regions overlaps the record assembly of this one
        t0 = time.perf_counter()
        records = calls.call_chromosome_streamed(table, clf.classify, options, genotype,
                                                 chunk_rows=getattr(options, "chunk_rows", 16384))
        t_calls = time.perf_counter() - t0
        calls.write_chromosome(out_path_prefix, records)
        logging.info("%s: %d rows: parse %.3f s, classify + calls %.3f s, write %.3f s -> %d records",
                     chrom or segments_out_file, len(table), t_parse, t_calls, time.perf_counter() - t0 - t_calls,
                     len(records))
        return len(records)
    if aggregate is None or write is None:
        raise ValueError("inject both of the reference's aggregate and write functions, or neither")
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
                 classifier=None, genotype=None):
        self.segments_out_file = segments_out_file
        self.chrom = chrom
        self.num_classes = 5
        self._aggregate, self._write, self._classifier, self._genotype = aggregate, write, classifier, genotype

    def run(self, out_path_prefix, options):
        return run_predict(self.segments_out_file, out_path_prefix, options, self._aggregate,
                           self._write, self._classifier, self.chrom, self._genotype)

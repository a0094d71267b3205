"""GPU parity AT THE BENCHED CONFIGURATION: one micro-batch of 10 000 sites (bench.py) and of 8192
(predict.get_classifier's default).  At B = 10 000 the largest activation buffers hold more than
2^31 elements, so the known-answer rows of tests/golden/cnn_golden.npz are placed at the head, in
the middle and in the LAST slots of the batch, where any 32-bit index arithmetic would go wrong.
Reference contract: src/network/predict.py:209-210 (labels = argmax, softmax) on the images of
src/network/create_batch.py:88-155.  Tolerances are the north-star's."""
import numpy as np
import pytest
import torch

from svision_b200 import classifier as C, sites

pytestmark = pytest.mark.gpu

SOFTMAX_TOL = 1e-3
LOGIT_TOL = 4e-3


def _batch_with_known_answers(n, golden_rows):
    """n P1 rows with the 256 golden rows at [0, 96), [n//2, n//2 + 96) and [n - 64, n)."""
    rows = sites.make_sites_p1(n, seed=sites.SEED_CONFIG2).copy()
    where = np.concatenate([np.arange(0, 96), np.arange(n // 2, n // 2 + 96), np.arange(n - 64, n)])
    rows[where] = golden_rows[:256]
    return rows, where


@pytest.mark.parametrize("max_batch", [10_000, 8192])
def test_known_answers_at_the_benched_micro_batch(max_batch, cnn_golden, synthetic_weights):
    ref_logits = torch.from_numpy(cnn_golden["logits_fp64"][:256])
    ref_labels = ref_logits.argmax(1).numpy().astype(np.int32)
    ref_probs = torch.softmax(ref_logits, 1).numpy()
    rows, where = _batch_with_known_answers(max_batch, cnn_golden["rows"])
    with C.Classifier(synthetic_weights, device=0, max_batch=max_batch) as clf:
        assert clf.max_batch == max_batch
        rd = clf.rows_to_device(rows)
        # device entry (what bench.py's `value` times)
        labels, probs, logits = clf.classify_device(rd, want_logits=True)
        labels, probs, logits = labels.cpu().numpy(), probs.cpu().numpy(), logits.cpu().numpy()
        assert np.array_equal(labels[where], ref_labels)
        assert np.abs(probs[where] - ref_probs).max() < SOFTMAX_TOL
        assert np.abs(logits[where] - ref_logits.numpy()).max() < LOGIT_TOL
        # every site of the full batch is a proper softmax row
        assert np.abs(probs.sum(1) - 1).max() < 1e-5
        assert np.array_equal(labels, probs.argmax(1).astype(np.int32))
        # host entry (what bench.py's `e2e` times) and the 8-byte calls entry: same bits
        l_host, p_host = clf.classify(rows)
        assert np.array_equal(l_host, labels) and np.array_equal(p_host, probs)
        call_l, call_s = clf.classify_device_calls(rd)
        assert np.array_equal(call_l.cpu().numpy(), labels)
        assert np.array_equal(call_s.cpu().numpy(), probs[np.arange(max_batch), labels])
        # a site's result does not depend on where in the micro-batch it sits: the same golden rows
        # in a small batch give the same bits as in the last slots of the full one
        l_small, p_small = clf.classify(rows[where])
        assert np.array_equal(l_small, labels[where]) and np.array_equal(p_small, probs[where])


def test_ragged_stream_over_full_micro_batches(cnn_golden, synthetic_weights):
    """2.3 micro-batches of 8192: the known answers sit in the ragged tail."""
    n = 8192 * 2 + 2500
    ref_logits = torch.from_numpy(cnn_golden["logits_fp64"][:256])
    rows, where = _batch_with_known_answers(n, cnn_golden["rows"])
    with C.Classifier(synthetic_weights, device=0, max_batch=8192) as clf:
        labels, probs = clf.classify(rows)
    assert np.array_equal(labels[where], ref_logits.argmax(1).numpy().astype(np.int32))
    assert np.abs(probs[where] - torch.softmax(ref_logits, 1).numpy()).max() < SOFTMAX_TOL

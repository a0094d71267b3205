"""Several GPUs: in one process behind the C-ABI (svx_multi_*, what the reference's single SVision
process would hold: SVision:296-341) and across processes (fused exchange under torchrun).  The
two-GPU cases skip on a single-GPU box; the driver's round-end run and profiles/ record them."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from svision_b200 import classifier as C, sites

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_handle_on_one_device_equals_single_handle(cnn_golden, synthetic_weights):
    """svx_multi with one device: chunking (3 chunks, the last ragged) must not change a bit."""
    rows = np.concatenate([cnn_golden["rows"], sites.make_sites_p1(5000, seed=sites.SEED_CONFIG3)])
    with C.Classifier(synthetic_weights, device=0, max_batch=2048) as one:
        ref_l, ref_p = one.classify(rows)
    with C.MultiClassifier(synthetic_weights, devices=[0], max_batch=2048) as multi:
        l, p = multi.classify(rows)
        assert multi.last_split() == [rows.shape[0]]
        l0, p0 = multi.classify(np.zeros((0, 12), np.int32))
        assert l0.shape == (0,) and p0.shape == (0, 5)
    assert np.array_equal(l, ref_l) and np.array_equal(p, ref_p)
    ref_logits = torch.from_numpy(cnn_golden["logits_fp64"])
    assert np.array_equal(l[:256], ref_logits.argmax(1).numpy().astype(np.int32))
    assert np.abs(p[:256] - torch.softmax(ref_logits, 1).numpy()).max() < 1e-3


def test_multi_handle_rejects_bad_device_lists(synthetic_weights):
    from svision_b200 import _lib
    with pytest.raises(_lib.SvxError):
        C.MultiClassifier(synthetic_weights, devices=[0, 0], max_batch=64)
    with pytest.raises(_lib.SvxError):
        C.MultiClassifier(synthetic_weights, devices=[torch.cuda.device_count()], max_batch=64)


def test_multi_handle_two_devices_in_one_process(cnn_golden, synthetic_weights):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rows = np.concatenate([cnn_golden["rows"], sites.make_sites_p1(30_000, seed=sites.SEED_CONFIG3)])
    with C.Classifier(synthetic_weights, device=0, max_batch=4096) as one:
        ref_l, ref_p = one.classify(rows)
    with C.MultiClassifier(synthetic_weights, devices=[0, 1], max_batch=4096) as multi:
        l, p = multi.classify(rows)
        split = multi.last_split()
        assert sum(split) == rows.shape[0] and min(split) > 0          # both devices took chunks
        l2, p2 = multi.classify(rows[:300])                            # small call: an equal split
        assert multi.last_split() == [150, 150]
    assert np.array_equal(l, ref_l) and np.array_equal(p, ref_p)       # device-independent bits
    assert np.array_equal(l2, ref_l[:300]) and np.array_equal(p2, ref_p[:300])


def _torchrun(script_args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])


def test_fused_exchange_two_gpus_bit_identical_and_timeout_detected():
    """tools/exchange_check.py under torchrun: fused exchange == NCCL all-gather == single-GPU result,
    bit for bit; a rank that shows up late is reported (sticky error, its calls poisoned), never
    answered with stale results, and the exchange recovers."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    res = _torchrun([os.path.join(ROOT, "tools", "exchange_check.py"), "--sites", "3000", "--iters", "3",
                     "--timeout-test"], 29533)
    assert res["world"] == 2 and res["all_paths_bit_identical"] is True
    assert res["timeout_test"].startswith("detected")


def test_bench_two_gpus_checks_every_rank_slice():
    """bench.py at N=2: the spot check covers the WHOLE gathered buffer (both ranks' known answers)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    res = _torchrun([os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3",
                     "--no-cpu-baseline"], 29534)
    assert res["n_gpus"] == 2
    assert res["parity_spot_check"]["known_answers_checked"] == 2 * 256
    assert res["parity_spot_check"]["ok"] is True
    assert res["strong_100k"]["parity_ok"] is True


def test_get_classifier_with_a_device_list(synthetic_weights, tmp_path, cnn_golden):
    """predict.get_classifier(devices=[...]): one device -> a plain Classifier; two -> ONE object driving both
    from this process (what `python -m svision_b200.step2 --devices all` uses), same bits."""
    from svision_b200 import predict, tf_bundle
    prefix = str(tmp_path / "m.ckpt")
    tf_bundle.write_bundle(prefix, synthetic_weights, data_crc=False)
    one = predict.get_classifier(prefix, devices=[0], max_batch=256)
    assert isinstance(one, C.Classifier)
    rows = cnn_golden["rows"]
    l1, p1 = one.classify(rows)
    if torch.cuda.device_count() >= 2:
        two = predict.get_classifier(prefix, devices=[0, 1], max_batch=256)
        assert isinstance(two, C.MultiClassifier)
        l2, p2 = two.classify(rows)
        assert np.array_equal(l1, l2) and np.array_equal(p1, p2) and min(two.last_split()) > 0
        two.close()
    one.close()
    predict._CLASSIFIER_CACHE.clear()

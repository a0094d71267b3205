"""GPU parity: tcgen05 layer kernel and the full encode+classify path against the oracle.
Tolerances are the north-star's: labels identical, softmax within 1e-3 (fp32)."""
import numpy as np
import pytest
import torch

from oracle import alexnet, encoder_c
from svision_b200 import classifier as C, sites

pytestmark = pytest.mark.gpu

SOFTMAX_TOL = 1e-3          # BASELINE.json north_star
LOGIT_TOL = 4e-3            # |dsoftmax| <= 1e-3  =>  |dlogit| <~ 4e-3


@pytest.mark.parametrize("m,n,k,bn", [(128, 128, 64, 128), (1000, 256, 512, 128), (300, 96, 192, 96),
                                      (700, 384, 320, 192), (513, 512, 1024, 256), (2048, 1024, 9216, 256)])
def test_gemm_kernel_3pass(m, n, k, bn):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda")
    b = torch.randn(n, k, device="cuda") * 0.05
    ref = a.double() @ b.double().T
    c = C.gemm_selftest(a, b, block_n=bn, precision="3pass")
    # hi/lo split carries ~22 bits; fp32 accumulation over k terms
    assert (c.double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 1e-5
    c1 = C.gemm_selftest(a, b, block_n=bn, precision="1pass")
    assert (c1.double() - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("grid_w,taps,center,k,n,bn", [
    (29, 5, 2, 128, 256, 128),      # conv2-like: 5x5 over a 29-wide grid (slab of 248 rows)
    (14, 3, 1, 256, 384, 192),      # conv3/4-like
    (14, 3, 1, 192, 256, 128),      # conv5-like
    (57, 3, 0, 64, 96, 96),         # dense conv1: 3x3 over the space-to-depth grid, 96-column tile
])
def test_shifted_gemm_kernel_matches_fp64(grid_w, taps, center, k, n, bn):
    """The layer kernel as the shifted GEMM it is: C[m] = sum_t A[m + off_t] @ B_t^T, rows outside A
    read as zero (TMA zero fill), against an fp64 evaluation of the same sum."""
    torch.manual_seed(grid_w * 100 + taps)
    m = 5 * grid_w * grid_w + 37
    offs = [(kh - center) * grid_w + (kw - center) for kh in range(taps) for kw in range(taps)]
    a = torch.randn(m, k, device="cuda")
    b = torch.randn(n, len(offs) * k, device="cuda") * 0.05
    c = C.conv_selftest(a, b, offs, block_n=bn, precision="3pass")
    ref = torch.zeros(m, n, dtype=torch.float64, device="cuda")
    ad = a.double()
    for t, off in enumerate(offs):
        shifted = torch.zeros_like(ad)
        lo, hi = max(0, -off), min(m, m - off)
        shifted[lo:hi] = ad[lo + off:hi + off]
        ref += shifted @ b[:, t * k:(t + 1) * k].double().T
    assert (c.double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 1e-5


@pytest.fixture(scope="module")
def clf(synthetic_weights):
    c = C.Classifier(synthetic_weights, device=0, max_batch=64)
    yield c
    c.close()


# conv2 / conv5 at full resolution are never materialised (their max-pool runs in the epilogue of the
# layer kernel): norm2 = LRN(pool(conv2)) and pool5 = pool(conv5) pin them
LAYERS = ("norm1", "norm2", "conv3", "conv4", "pool5", "fc6", "fc7")


def _check_layers(clf, inter, n, names):
    for name in names:
        got = clf.debug_activation(name, n)
        ref = inter[name].numpy()
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() < 2e-4 * scale + 1e-5, name


def test_layerwise_activations_match_oracle(clf, cnn_golden, synthetic_weights):
    rows = cnn_golden["rows"][:16]
    imgs = encoder_c.encode_f32(rows)
    _, inter = alexnet.forward(imgs, synthetic_weights, torch.float32, return_intermediates=True)
    # classify path: fused sparse front end (rows -> norm1) + tensor-core layers
    clf.classify_device(clf.rows_to_device(rows))
    torch.cuda.synchronize()
    _check_layers(clf, inter, rows.shape[0], LAYERS)
    # forward path on materialised images: dense tcgen05 conv1 + pool1/LRN1 kernels
    clf.forward(torch.from_numpy(imgs).cuda())
    torch.cuda.synchronize()
    _check_layers(clf, inter, rows.shape[0], ("conv1",) + LAYERS)


def test_fused_front_end_matches_dense_path(clf, cnn_golden):
    """classify (sparse fused front end) vs forward on the materialised images (dense tcgen05 conv1 +
    pool1/LRN1): the same maths with the fp32 terms in a different order."""
    rows = np.concatenate([sites.edge_case_sites(), cnn_golden["rows"][:50]])
    rd = clf.rows_to_device(rows)
    labels, probs, logits = clf.classify_device(rd, want_logits=True)
    torch.cuda.synchronize()
    norm1_fused = clf.debug_activation("norm1", rows.shape[0])
    logits_dense = clf.forward(clf.encode(rd, dtype=torch.float16))
    torch.cuda.synchronize()
    norm1_dense = clf.debug_activation("norm1", rows.shape[0])
    assert np.abs(norm1_fused - norm1_dense).max() < 1e-3          # values up to ~140
    assert (logits - logits_dense).abs().max().item() < 1e-3
    assert np.array_equal(labels.cpu().numpy(), logits_dense.argmax(1).cpu().numpy().astype(np.int32))
    assert (probs - torch.softmax(logits_dense, 1)).abs().max().item() < 2e-4


def test_labels_and_softmax_match_oracle(clf, cnn_golden):
    rows = cnn_golden["rows"]
    ref_logits = torch.from_numpy(cnn_golden["logits_fp64"])
    labels, probs, logits = clf.classify_device(clf.rows_to_device(rows), want_logits=True)
    assert (logits.cpu().double() - ref_logits).abs().max().item() < LOGIT_TOL
    assert (probs.cpu().double() - torch.softmax(ref_logits, 1)).abs().max().item() < SOFTMAX_TOL
    assert np.array_equal(labels.cpu().numpy(), ref_logits.argmax(1).numpy().astype(np.int32))


def test_host_entry_equals_device_entry_and_ragged_batches(clf, cnn_golden):
    # 256 rows through max_batch=64 -> 4 micro-batches; 77 rows -> ragged tail
    rows = cnn_golden["rows"]
    l_dev, p_dev = clf.classify_device(clf.rows_to_device(rows))
    l_host, p_host = clf.classify(rows)
    assert l_host.dtype == np.int32 and p_host.dtype == np.float32
    assert np.array_equal(l_host, l_dev.cpu().numpy())
    assert np.array_equal(p_host, p_dev.cpu().numpy())
    l77, p77 = clf.classify(rows[:77])
    assert np.array_equal(l77, l_host[:77]) and np.array_equal(p77, p_host[:77])
    l0, p0 = clf.classify(np.zeros((0, 12), np.int32))
    assert l0.shape == (0,) and p0.shape == (0, 5)


def test_forward_on_encoder_images_equals_fused_path(clf, cnn_golden):
    rows = cnn_golden["rows"][:48]
    rd = clf.rows_to_device(rows)
    _, _, logits = clf.classify_device(rd, want_logits=True)
    l16 = clf.forward(clf.encode(rd, dtype=torch.float16))
    l32 = clf.forward(clf.encode(rd, dtype=torch.float32))
    assert torch.equal(l16, l32)                       # both image dtypes are lossless
    # dense conv1 (forward) vs fused sparse front end (classify): same maths, different fp32 order
    assert (l16 - logits).abs().max().item() < 1e-3


def test_results_independent_of_batch_position(clf, cnn_golden):
    # sites are independent: permuting the batch permutes the results bit-for-bit
    rows = cnn_golden["rows"][:64]
    perm = np.random.default_rng(0).permutation(64)
    l1, p1 = clf.classify(rows)
    l2, p2 = clf.classify(rows[perm])
    assert np.array_equal(l1[perm], l2) and np.array_equal(p1[perm], p2)


def test_softmax_rows_sum_to_one_full_size(clf):
    rows = sites.make_sites_p1(2000, seed=sites.SEED_CONFIG2)
    labels, probs = clf.classify(rows)
    assert np.abs(probs.sum(1) - 1).max() < 1e-5
    assert ((labels >= 0) & (labels < 5)).all()
    assert np.array_equal(labels, probs.argmax(1).astype(np.int32))


def test_pooled_epilogue_across_warp_and_tile_boundaries(clf, cnn_golden, synthetic_weights):
    """norm2 / pool5 come from window maxima that the conv2 / conv5 epilogues accumulate across warps,
    CTAs and tiles (red.max): check EVERY site of a batch whose size is not a multiple of anything,
    twice (the pooled buffers must be back to zero after each pass)."""
    rows = np.concatenate([cnn_golden["rows"][:37], sites.edge_case_sites(), cnn_golden["rows"][200:213]])
    imgs = encoder_c.encode_f32(rows)
    _, inter = alexnet.forward(imgs, synthetic_weights, torch.float32, return_intermediates=True)
    rd = clf.rows_to_device(rows)
    for _ in range(2):
        clf.classify_device(rd)
        torch.cuda.synchronize()
        _check_layers(clf, inter, rows.shape[0], ("norm2", "pool5"))


def test_full_size_config2_properties(clf):
    """BASELINE config 2 size (10 000 sites): determinism, shard-independence and agreement of a
    strided sample with the CPU oracle (labels equal, softmax within 1e-3)."""
    rows = sites.make_sites_p1(10_000, seed=sites.SEED_CONFIG2)
    l1, p1 = clf.classify(rows)
    l2, p2 = clf.classify(rows)
    assert np.array_equal(l1, l2) and np.array_equal(p1, p2)                  # deterministic
    lo, po = clf.classify(rows[5000:])                                         # any contiguous shard
    assert np.array_equal(lo, l1[5000:]) and np.array_equal(po, p1[5000:])
    assert np.abs(p1.sum(1) - 1).max() < 1e-5
    assert len(np.unique(l1)) >= 3                                             # calibrated weights: classes vary
    idx = np.arange(0, 10_000, 157)
    from svision_b200 import weights as W
    ref_l, ref_p, _ = alexnet.classify(encoder_c.encode_f32(rows[idx]), W.synthetic_weights(), torch.float32, batch=64)
    assert np.array_equal(l1[idx], ref_l.astype(np.int32))
    assert np.abs(p1[idx] - ref_p).max() < SOFTMAX_TOL


def test_one_pass_mode_is_measurably_worse(cnn_golden, synthetic_weights):
    """The 1-pass mode exists for comparison only: it must NOT be mistaken for parity-clean."""
    rows = cnn_golden["rows"][:64]
    ref_logits = torch.from_numpy(cnn_golden["logits_fp64"][:64])
    with C.Classifier(synthetic_weights, device=0, max_batch=64, precision="1pass") as c:
        _, probs, _ = c.classify_device(c.rows_to_device(rows), want_logits=True)
    err = (probs.cpu().double() - torch.softmax(ref_logits, 1)).abs().max().item()
    assert err > SOFTMAX_TOL


def test_real_demo_rows_and_ont_profile_match_oracle(clf, synthetic_weights):
    """Real-data rows (reference demo BAM) and ONT-profile rows (BASELINE config 5 shape):
    labels equal the CPU oracle, softmax within 1e-3."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "demo_rows_golden.npz"))
    rows = np.concatenate([g["rows"], sites.make_sites_p1(128, seed=sites.SEED_CONFIG5, profile="ont")])
    labels, probs = clf.classify(rows)
    ref_l, ref_p, _ = alexnet.classify(encoder_c.encode_f32(rows), synthetic_weights, torch.float32, batch=100)
    assert np.array_equal(labels, ref_l.astype(np.int32))
    assert np.abs(probs - ref_p).max() < SOFTMAX_TOL


def test_gpu_matches_the_independent_numpy_oracle(clf, synthetic_weights):
    """The second oracle (numpy fp64 from the TF op definitions, no torch) on fresh rows: labels equal,
    softmax within 1e-3, layerwise activations of the pooled layers within 2e-4."""
    from oracle import alexnet_np
    rows = np.concatenate([sites.make_sites_p2(20, seed=99), sites.make_sites_p1(12, seed=5, profile="ont")])
    logits, inter = alexnet_np.forward(encoder_c.encode_f32(rows), synthetic_weights, return_intermediates=True)
    labels, probs, got_logits = clf.classify_device(clf.rows_to_device(rows), want_logits=True)
    torch.cuda.synchronize()
    assert np.array_equal(labels.cpu().numpy(), alexnet_np.argmax_first(logits).astype(np.int32))
    assert np.abs(probs.cpu().numpy() - alexnet_np.softmax(logits)).max() < SOFTMAX_TOL
    assert np.abs(got_logits.cpu().numpy() - logits).max() < LOGIT_TOL
    for name in ("norm1", "norm2", "pool5", "fc7"):
        got, ref = clf.debug_activation(name, rows.shape[0]), inter[name]
        assert np.abs(got - ref).max() < 2e-4 * np.abs(ref).max() + 1e-5, name


@pytest.mark.parametrize("max_batch", [1, 3, 129])
def test_odd_micro_batches_match_oracle(max_batch, cnn_golden, synthetic_weights):
    """Micro-batches far from any tile size (one image = 841 conv2 rows = 3.3 pair tiles; fewer tiles than
    CTA pairs; pooling windows of the last image cut by the last chunk): same answers."""
    rows = cnn_golden["rows"][:37]
    ref_logits = torch.from_numpy(cnn_golden["logits_fp64"][:37])
    with C.Classifier(synthetic_weights, device=0, max_batch=max_batch) as c:
        labels, probs = c.classify(rows)
    assert np.array_equal(labels, ref_logits.argmax(1).numpy().astype(np.int32))
    assert np.abs(probs - torch.softmax(ref_logits, 1).numpy()).max() < SOFTMAX_TOL

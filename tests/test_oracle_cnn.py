"""CPU: the torch AlexNet oracle against its committed golden logits, plus structural checks of
the restatement (shapes, LRN formula, grouped conv) against an independent numpy evaluation."""
import numpy as np
import torch

from oracle import alexnet, encoder_c
from svision_b200 import weights as W


def test_parameter_count():
    assert W.N_PARAMS == 56_888_709            # SURVEY §8(a) layer table


def test_oracle_matches_golden_logits(cnn_golden, synthetic_weights):
    rows = cnn_golden["rows"][:16]
    imgs = encoder_c.encode_f32(rows)
    torch.set_num_threads(4)
    l32 = alexnet.forward(imgs, synthetic_weights, torch.float32).numpy()
    assert np.abs(l32 - cnn_golden["logits_fp32"][:16]).max() < 1e-3
    assert np.abs(l32 - cnn_golden["logits_fp64"][:16]).max() < 2e-3
    assert (l32.argmax(1) == cnn_golden["logits_fp64"][:16].argmax(1)).all()


def test_golden_labels_are_balanced(cnn_golden):
    hist = np.bincount(cnn_golden["logits_fp64"].argmax(1), minlength=5)
    assert (hist > 0).all(), hist            # label parity is only meaningful if classes vary


def test_lrn_matches_tf_formula():
    # tf.nn.local_response_normalization(depth_radius=2, alpha=2e-5, beta=.75, bias=1)
    x = torch.rand(2, 96, 5, 5, dtype=torch.float64) * 50
    got = alexnet._lrn(x).numpy()
    xn = x.numpy()
    sq = np.pad(xn * xn, ((0, 0), (2, 2), (0, 0), (0, 0)))
    s = sum(sq[:, k:k + 96] for k in range(5))
    ref = xn / (1.0 + 2e-5 * s) ** 0.75
    assert np.abs(got - ref).max() < 1e-12


def test_grouped_conv_matches_split_concat(synthetic_weights):
    # alexnet.py:124-129: split input on channels, weights on the OUTPUT axis, concat
    x = torch.rand(1, 96, 9, 9, dtype=torch.float64)
    w = torch.from_numpy(synthetic_weights["conv2/weights"]).double()     # [5,5,48,256]
    b = torch.from_numpy(synthetic_weights["conv2/biases"]).double()
    got = alexnet._conv(x, synthetic_weights, "conv2", torch.float64, padding=2)
    outs = []
    for g in range(2):
        wg = w[..., g * 128:(g + 1) * 128].permute(3, 2, 0, 1)
        outs.append(torch.nn.functional.conv2d(x[:, g * 48:(g + 1) * 48], wg, padding=2))
    ref = torch.relu(torch.cat(outs, 1) + b[None, :, None, None])
    assert (got - ref).abs().max() < 1e-12


# ---- second, independent oracle (oracle/alexnet_np.py: numpy fp64 from the TF op definitions) ----------
def test_numpy_oracle_matches_torch_oracle_on_golden_rows(cnn_golden, synthetic_weights):
    """Two restatements that share no code (torch functional ops vs explicit im2col / SAME-pad /
    LRN / xw_plus_b arithmetic in numpy) must agree to fp64 round-off on the golden rows."""
    from oracle import alexnet_np
    rows = cnn_golden["rows"][:24]
    imgs = encoder_c.encode_f32(rows)
    logits, inter_np = alexnet_np.forward(imgs, synthetic_weights, return_intermediates=True)
    assert np.abs(logits - cnn_golden["logits_fp64"][:24]).max() < 1e-9
    assert np.array_equal(alexnet_np.argmax_first(logits), cnn_golden["logits_fp64"][:24].argmax(1))
    _, inter_t = alexnet.forward(imgs[:4], synthetic_weights, torch.float64, return_intermediates=True)
    for name in ("norm1", "norm2", "conv3", "conv4", "pool5", "fc6", "fc7"):
        assert np.abs(inter_np[name][:4] - inter_t[name].numpy()).max() < 1e-9, name
    p = alexnet_np.softmax(logits)
    assert np.abs(p - torch.softmax(torch.from_numpy(logits), 1).numpy()).max() < 1e-15


def test_numpy_oracle_same_padding_arithmetic():
    """TF SAME: out = ceil(in / s), pad_total = max((out-1)*s + k - in, 0), extra pixel at the end."""
    from oracle import alexnet_np
    assert alexnet_np._same_pads(27, 5, 1) == (27, 2, 2)
    assert alexnet_np._same_pads(13, 3, 1) == (13, 1, 1)
    assert alexnet_np._same_pads(10, 3, 2) == (5, 0, 1)          # asymmetric: the extra pad goes last
    assert alexnet_np._same_pads(227, 11, 4) == (57, 4, 4)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((1, 7, 6, 2))
    w = rng.standard_normal((3, 3, 2, 4))
    got = alexnet_np.conv2d(x, w, 2, "SAME")                      # out 4 x 3, pads (1,1) rows / (0,1) cols
    assert got.shape == (1, 4, 3, 4)
    xp = np.pad(x, ((0, 0), (1, 1), (0, 1), (0, 0)))
    ref = np.zeros((4, 3, 4))
    for oy in range(4):
        for ox in range(3):
            ref[oy, ox] = np.einsum("abc,abco->o", xp[0, 2 * oy:2 * oy + 3, 2 * ox:2 * ox + 3], w)
    assert np.abs(got[0] - ref).max() < 1e-12
    v = alexnet_np.conv2d(x, w, 2, "VALID")
    assert v.shape == (1, 3, 2, 4)


def test_numpy_oracle_argmax_takes_first_maximum():
    from oracle import alexnet_np
    assert alexnet_np.argmax_first(np.array([[1.0, 3.0, 3.0, 0.0, 3.0]]))[0] == 1

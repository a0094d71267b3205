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

"""CPU: TF-free checkpoint (tensor bundle) reader.  No real SVision checkpoint exists here, so this
is a writer<->reader round trip plus known-answer checks of the building blocks (crc32c test
vectors from RFC 3720, leveldb masking, footer magic)."""
import struct

import numpy as np
import pytest

from svision_b200 import tf_bundle, weights


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert tf_bundle.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tf_bundle.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tf_bundle.crc32c(bytes(range(32))) == 0x46DD794E
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283
    big = bytes(range(256)) * 9 + b"xyz"          # exercises the 8-byte-sliced path and the tail
    ref = 0
    ref = tf_bundle.crc32c(big[:100])
    assert tf_bundle.crc32c(big) == tf_bundle.crc32c(big[100:], ref)
    assert tf_bundle.mask_crc(0) == 0xA282EAD8


def _small_model(rng):
    return {"conv1/weights": rng.standard_normal((3, 3, 2, 4), dtype=np.float32),
            "conv1/biases": rng.standard_normal(4, dtype=np.float32),
            "fc8/weights": rng.standard_normal((7, 5), dtype=np.float32),
            "fc8/biases": rng.standard_normal(5, dtype=np.float32),
            "global_step_as_float": np.array(3.0, dtype=np.float32)}


def test_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    model = _small_model(rng)
    prefix = str(tmp_path / "m.ckpt")
    tf_bundle.write_bundle(prefix, model)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == tf_bundle.TABLE_MAGIC
    index = tf_bundle.read_index(prefix)
    assert index[""]["num_shards"] == 1
    assert index["conv1/weights"]["shape"] == (3, 3, 2, 4) and index["conv1/weights"]["dtype"] == 1
    got = tf_bundle.read_bundle(prefix, verify_data=True)
    assert set(got) == set(model)
    for k in model:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], model[k])
    only = tf_bundle.read_bundle(prefix, names=["fc8/biases"])
    assert list(only) == ["fc8/biases"]


def test_bundle_many_keys_prefix_compression(tmp_path):
    # > 16 keys with long common prefixes: restart points and shared-prefix decoding
    rng = np.random.default_rng(1)
    model = {f"layer_with_a_long_name/{i:03d}/weights": rng.standard_normal(i + 1, dtype=np.float32)
             for i in range(40)}
    prefix = str(tmp_path / "many.ckpt")
    tf_bundle.write_bundle(prefix, model)
    got = tf_bundle.read_bundle(prefix, verify_data=True)
    assert all(np.array_equal(got[k], model[k]) for k in model)


def test_bundle_failures_are_loud(tmp_path):
    rng = np.random.default_rng(2)
    prefix = str(tmp_path / "m.ckpt")
    tf_bundle.write_bundle(prefix, _small_model(rng))
    with pytest.raises(KeyError):
        tf_bundle.read_bundle(prefix, names=["conv9/weights"])
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[5] ^= 0x40                                  # corrupt the data block
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tf_bundle.read_index(prefix)
    raw = bytearray(open(prefix + ".index", "rb").read())
    raw[-1] ^= 0xFF                                 # corrupt the magic
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tf_bundle.read_index(prefix)
    # flipped payload byte is caught when data verification is on
    prefix2 = str(tmp_path / "n.ckpt")
    tf_bundle.write_bundle(prefix2, _small_model(rng))
    d = bytearray(open(prefix2 + ".data-00000-of-00001", "rb").read())
    d[0] ^= 1
    open(prefix2 + ".data-00000-of-00001", "wb").write(bytes(d))
    with pytest.raises(ValueError):
        tf_bundle.read_bundle(prefix2, verify_data=True)


def test_load_checkpoint_checks_the_variable_set(tmp_path):
    rng = np.random.default_rng(3)
    # a structurally complete (tiny-valued) model in the reference's 16 variables
    model = {}
    for layer, shape in weights.WEIGHT_SHAPES.items():
        if layer in ("fc6", "fc7"):
            model[f"{layer}/weights"] = np.zeros(shape, dtype=np.float32)
        else:
            model[f"{layer}/weights"] = rng.standard_normal(shape, dtype=np.float32)
        model[f"{layer}/biases"] = rng.standard_normal(shape[-1], dtype=np.float32)
    prefix = str(tmp_path / "svision-cnn-model.ckpt")
    tf_bundle.write_bundle(prefix, model, data_crc=False)
    got = weights.load_checkpoint(prefix)
    assert set(got) == set(weights.VARIABLE_NAMES)
    assert np.array_equal(got["conv2/weights"], model["conv2/weights"])
    del model["fc7/biases"]
    tf_bundle.write_bundle(prefix, model, data_crc=False)
    with pytest.raises(KeyError):
        weights.load_checkpoint(prefix)

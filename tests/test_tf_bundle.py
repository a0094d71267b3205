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


# ---- hand-built indexes shaped like what tf.train.Saver really writes -------------------------------------
# (the writer in tf_bundle.py emits one data block and only float entries; a real TF 1.x save of a trained
#  model also holds optimizer slots, int64 counters, several data blocks and possibly several shards)
def _entry(dtype: int, shape, shard: int, offset: int, size: int, crc: int = 0, sliced: bool = False) -> bytes:
    pv = tf_bundle._put_varint
    dims = b"".join(b"\x12" + pv(len(d)) + d for d in (b"\x08" + pv(int(s)) for s in shape))
    out = b"\x08" + pv(dtype) + b"\x12" + pv(len(dims)) + dims
    if shard:
        out += b"\x18" + pv(shard)
    out += b"\x20" + pv(offset) + b"\x28" + pv(size) + b"\x35" + struct.pack("<I", crc)
    if sliced:
        out += b"\x3a\x00"                               # field 7 (slices), empty TensorSliceProto
    return out


def _write_index(path, items, per_block=4, header=b"\x08\x01\x10\x00\x1a\x02\x08\x01", tag=0):
    """items: sorted [(key bytes, value bytes)]; `per_block` entries per data block, as many index
    entries as blocks (a multi-block table); `tag` = compression byte written after every block."""
    items = [(b"", header)] + list(items)
    out = bytearray()

    def emit(block: bytes):
        off = len(out)
        out.extend(block)
        out.append(tag)
        out.extend(struct.pack("<I", tf_bundle.mask_crc(tf_bundle.crc32c(block + bytes([tag])))))
        return off, len(block)

    handles = []
    for i in range(0, len(items), per_block):
        chunk = items[i:i + per_block]
        off, size = emit(tf_bundle._build_block(chunk))
        handles.append((chunk[-1][0] + b"\x00", tf_bundle._put_varint(off) + tf_bundle._put_varint(size)))
    m_off, m_size = emit(tf_bundle._build_block([]))
    i_off, i_size = emit(tf_bundle._build_block(handles))
    pv = tf_bundle._put_varint
    footer = pv(m_off) + pv(m_size) + pv(i_off) + pv(i_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", tf_bundle.TABLE_MAGIC)
    out.extend(footer)
    open(path, "wb").write(bytes(out))


_DATA_CACHE = {}


def _trained_style_checkpoint(tmp_path, shards=1, **index_kw):
    """The reference's 16 variables plus what an Adam-trained TF 1.x graph saves beside them.  The
    228 MB of tensor data are written once per shard count; each test gets its own index."""
    if shards not in _DATA_CACHE:
        import tempfile
        rng = np.random.default_rng(11)
        tensors = {}
        for layer, shape in weights.WEIGHT_SHAPES.items():
            small = layer in ("fc6", "fc7")
            wt = np.zeros(shape, np.float32) if small else rng.standard_normal(shape, dtype=np.float32)
            tensors[f"{layer}/weights"] = wt
            tensors[f"{layer}/biases"] = rng.standard_normal(shape[-1], dtype=np.float32)
            for slot in ("Adam", "Adam_1"):                    # optimizer slots: same shape, float
                tensors[f"{layer}/biases/{slot}"] = rng.standard_normal(shape[-1], dtype=np.float32)
        tensors["beta1_power"] = np.array(0.9, np.float32)
        tensors["beta2_power"] = np.array(0.999, np.float32)
        tensors["global_step"] = np.array(12345, np.int64)     # DT_INT64 = 9
        base = tempfile.mkdtemp(prefix="svx_ckpt_")
        prefix = base + "/svision-cnn-model.ckpt"
        files = [open(f"{prefix}.data-{s:05d}-of-{shards:05d}", "wb") for s in range(shards)]
        offsets = [0] * shards
        items = []
        for k, name in enumerate(sorted(tensors)):
            arr = tensors[name]
            raw = arr.astype("<i8" if arr.dtype == np.int64 else "<f4").tobytes()
            s = k % shards
            files[s].write(raw)
            crc = tf_bundle.mask_crc(tf_bundle.crc32c(raw)) if len(raw) < 1 << 16 else 0
            items.append((name.encode(), _entry(9 if arr.dtype == np.int64 else 1, arr.shape, s, offsets[s], len(raw), crc)))
            offsets[s] += len(raw)
        for f in files:
            f.close()
        _DATA_CACHE[shards] = (prefix, tensors, items)
    prefix, tensors, items = _DATA_CACHE[shards]
    header = b"\x08" + tf_bundle._put_varint(shards) + b"\x10\x00\x1a\x02\x08\x01"
    _write_index(prefix + ".index", items, header=header, **index_kw)
    return prefix, tensors


def test_trained_style_checkpoint_multi_block_index_and_extra_keys(tmp_path):
    prefix, tensors = _trained_style_checkpoint(tmp_path, per_block=5)
    index = tf_bundle.read_index(prefix)
    assert len(index) == 1 + len(tensors)                                  # header + every key of 11 blocks
    assert index["global_step"]["dtype"] == 9 and index["global_step"]["shape"] == ()
    got = weights.load_checkpoint(prefix)                                  # exactly the 16 variables
    assert set(got) == set(weights.VARIABLE_NAMES)
    for k in weights.VARIABLE_NAMES:
        assert np.array_equal(got[k], tensors[k])
    everything = tf_bundle.read_bundle(prefix, verify_data=False)          # non-float entries are skipped
    assert "global_step" not in everything and "conv1/biases/Adam_1" in everything
    with pytest.raises(TypeError):
        tf_bundle.read_bundle(prefix, names=["global_step"])               # asked for by name: loud


def test_trained_style_checkpoint_two_shards(tmp_path):
    prefix, tensors = _trained_style_checkpoint(tmp_path, shards=2, per_block=7)
    assert tf_bundle.read_index(prefix)[""]["num_shards"] == 2
    got = weights.load_checkpoint(prefix)
    assert all(np.array_equal(got[k], tensors[k]) for k in weights.VARIABLE_NAMES)


def test_snappy_tagged_blocks_fail_loudly(tmp_path):
    prefix, _ = _trained_style_checkpoint(tmp_path, per_block=5, tag=1)    # tag 1 = snappy in leveldb tables
    with pytest.raises(ValueError, match="compressed"):
        weights.load_checkpoint(prefix)


def test_sliced_and_big_endian_and_truncated_fail_loudly(tmp_path):
    rng = np.random.default_rng(5)
    raw = rng.standard_normal(4, dtype=np.float32).tobytes()
    prefix = str(tmp_path / "s.ckpt")
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    _write_index(prefix + ".index", [(b"v", _entry(1, (4,), 0, 0, 16, sliced=True))])
    with pytest.raises(ValueError, match="slices"):
        tf_bundle.read_index(prefix)
    _write_index(prefix + ".index", [(b"v", _entry(1, (4,), 0, 0, 16))], header=b"\x08\x01\x10\x01")
    with pytest.raises(ValueError, match="big-endian"):
        tf_bundle.read_index(prefix)
    _write_index(prefix + ".index", [(b"v", _entry(1, (5,), 0, 0, 16))])   # 16 bytes for 5 floats
    with pytest.raises(ValueError, match="bytes for shape"):
        tf_bundle.read_bundle(prefix)
    _write_index(prefix + ".index", [(b"v", _entry(1, (8,), 0, 0, 32))])   # entry runs past the data file
    with pytest.raises(ValueError):
        tf_bundle.read_bundle(prefix)
    open(prefix + ".index", "wb").write(b"\x00" * 20)
    with pytest.raises(ValueError, match="too short"):
        tf_bundle.read_index(prefix)


def test_wrong_shape_in_checkpoint_is_rejected(tmp_path):
    rng = np.random.default_rng(6)
    model = {}
    for layer, shape in weights.WEIGHT_SHAPES.items():
        model[f"{layer}/weights"] = np.zeros(shape, np.float32)
        model[f"{layer}/biases"] = rng.standard_normal(shape[-1], dtype=np.float32)
    model["fc8/weights"] = np.zeros((4096, 6), np.float32)                 # a 6-class head
    prefix = str(tmp_path / "w.ckpt")
    tf_bundle.write_bundle(prefix, model, data_crc=False)
    with pytest.raises(ValueError, match="fc8/weights"):
        weights.load_checkpoint(prefix)

"""Checkpoints in the reference's format and naming (biggan_tensorflow_b200/checkpoint.py; saver: BigGAN.py:1018,
:1255-1284; names: SURVEY A.3).  Host-only: CRC-32C known answers, the table / bundle containers (including an index
file assembled by hand from the format description, so the reader is not only tested against our own writer),
corruption detection, the swapping-saver naming."""
import os
import struct

import numpy as np
import pytest

try:
    from biggan_tensorflow_b200 import checkpoint as ck
except RuntimeError:                       # libbiggan_b200.so not built yet (fresh checkout): build, then import
    import __graft_entry__
    __graft_entry__.build()
    from biggan_tensorflow_b200 import checkpoint as ck


def test_crc32c_known_answers_and_mask():
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA                     # RFC 3720 B.4
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(b"6789", ck.crc32c(b"12345")) == 0xE3069283  # streaming
    assert ck.crc32c(b"") == 0
    assert ck.mask_crc(0) == 0xA282EAD8
    c = 0xE3069283
    assert ck.mask_crc(c) == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_header_and_entry_protos():
    assert ck.encode_header(1) == b"\x08\x01\x1a\x02\x08\x01"
    e = ck.encode_entry(ck.DT_FLOAT, (3, 3, 96, 192), 1024, 663552, 0xDEADBEEF)
    assert e == (b"\x08\x01" + b"\x12\x11" + b"\x12\x02\x08\x03" + b"\x12\x02\x08\x03" + b"\x12\x02\x08\x60" + b"\x12\x03\x08\xc0\x01"
                 + b"\x20\x80\x08" + b"\x28\x80\xc0\x28" + b"\x35" + struct.pack("<I", 0xDEADBEEF))
    d = ck.decode_entry(e)
    assert d["dtype"] == 1 and d["shape"] == [3, 3, 96, 192] and d["offset"] == 1024 and d["size"] == 663552 and d["crc32c"] == 0xDEADBEEF
    scalar = ck.decode_entry(ck.encode_entry(ck.DT_FLOAT, (), 0, 4, 1))
    assert scalar["shape"] == [] and scalar["offset"] == 0 and scalar["size"] == 4


def _hand_made_index(entries):
    """A one-block table written byte by byte from the format description (no code shared with build_table)."""
    block = b""
    for k, v in entries:                                           # no prefix sharing: shared = 0 is always legal
        block += bytes([0, len(k), len(v)]) + k + v
    block += struct.pack("<I", 0) + struct.pack("<I", 1)          # one restart point at offset 0
    def trailer(b):
        return b"\x00" + struct.pack("<I", ck.mask_crc(ck.crc32c(b + b"\x00")))
    out = block + trailer(block)
    meta = struct.pack("<I", 0) + struct.pack("<I", 1)
    meta_off = len(out)
    out += meta + trailer(meta)
    handle = bytes([0, len(block)])                               # varint offset 0, varint size (< 128)
    idx = bytes([0, 1, len(handle)]) + b"\x7f" + handle + struct.pack("<I", 0) + struct.pack("<I", 1)
    idx_off = len(out)
    out += idx + trailer(idx)
    footer = bytes([meta_off, len(meta)]) + bytes([idx_off, len(idx)])
    footer += b"\x00" * (40 - len(footer)) + bytes.fromhex("57fb808b247547db")
    return out + footer


def test_reader_parses_a_hand_assembled_checkpoint(tmp_path):
    data = np.arange(6, dtype="<f4").tobytes()
    entries = [(b"", b"\x08\x01\x1a\x02\x08\x01"),
               (b"g/u", b"\x08\x01\x12\x08\x12\x02\x08\x01\x12\x02\x08\x06\x28\x18\x35" + struct.pack("<I", ck.mask_crc(ck.crc32c(data))))]
    idx = _hand_made_index(entries)
    assert len(idx) < 128 + 48 + 64
    prefix = str(tmp_path / "m")
    open(prefix + ".index", "wb").write(idx)
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    t = ck.read_bundle(prefix)
    assert list(t) == ["g/u"] and t["g/u"].shape == (1, 6) and t["g/u"].tolist() == [[0, 1, 2, 3, 4, 5]]


def test_table_round_trip_with_prefix_compression_restarts_and_many_blocks(monkeypatch):
    keys = sorted(("generator/resblock_up_%d/res%d/deconv_0/kernel/%s" % (i, j, s)).encode() for i in range(40) for j in (1, 2)
                  for s in ("", "Adam", "Adam_1", "ExponentialMovingAverage"))
    items = [(b"", b"hdr")] + [(k, k[::-1] * 3) for k in keys]
    for block_size in (ck.BLOCK_SIZE, 2000):
        monkeypatch.setattr(ck, "BLOCK_SIZE", block_size)
        buf = ck.build_table(items)
        assert ck.parse_table(buf) == items
        assert struct.unpack("<Q", buf[-8:])[0] == 0xDB4775248B80FB57
    bad = bytearray(buf)
    bad[10] ^= 1
    with pytest.raises(ValueError, match="checksum"):
        ck.parse_table(bytes(bad))
    with pytest.raises(ValueError, match="magic"):
        ck.parse_table(buf[:-1] + b"\x00")


def test_bundle_round_trip_and_corruption(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"generator/dense1/kernel": rng.standard_normal((34, 64)).astype(np.float32),
               "generator/dense1/u": rng.standard_normal((1, 64)).astype(np.float32),
               "beta1_power": np.array(0.0, np.float32), "beta2_power": np.array(0.81, np.float32),
               "global_step": np.array(7, np.int64), "discriminator/D_logit/bias": np.zeros(1, np.float32)}
    prefix = str(tmp_path / "sub" / "BigGAN.model-7")
    ck.write_bundle(prefix, tensors)
    back = ck.read_bundle(prefix)
    assert sorted(back) == sorted(tensors)
    for k in tensors:
        assert back[k].dtype == tensors[k].dtype and back[k].shape == tensors[k].shape and np.array_equal(back[k], tensors[k])
    assert list(ck.read_bundle(prefix, names={"beta2_power"})) == ["beta2_power"]
    ck.write_checkpoint_state(str(tmp_path / "sub"), "BigGAN.model-7")
    assert ck.latest_checkpoint(str(tmp_path / "sub")) == prefix
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[5] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    with pytest.raises(ValueError, match="checksum mismatch"):
        ck.read_bundle(prefix)
    assert len(ck.read_bundle(prefix, verify=False)) == len(tensors)


def test_swapping_saver_naming_round_trip():
    rng = np.random.default_rng(1)
    g_names, d_names = ["generator/G_logit/kernel", "generator/batch_norm/gamma"], ["discriminator/D_logit/kernel"]
    variables = {n: rng.standard_normal((2, 3)).astype(np.float32) for n in g_names + d_names}
    variables.update({"generator/G_logit/u": rng.standard_normal((1, 3)).astype(np.float32),
                      "generator/batch_norm/moving_mean": np.zeros(3, np.float32)})
    adam = {n: (rng.standard_normal((2, 3)).astype(np.float32), rng.random((2, 3)).astype(np.float32)) for n in g_names + d_names}
    ema = {n: rng.standard_normal((2, 3)).astype(np.float32) for n in g_names}
    t = ck.to_tf_tensors(variables, g_names, d_names, adam, adam, ema, t_g=5, t_d=6, beta1=0.0, beta2=0.9)
    # a generator variable's own name carries the AVERAGE, the raw value sits under .../ExponentialMovingAverage
    assert np.array_equal(t["generator/G_logit/kernel"], ema["generator/G_logit/kernel"])
    assert np.array_equal(t["generator/G_logit/kernel/ExponentialMovingAverage"], variables["generator/G_logit/kernel"])
    assert np.array_equal(t["discriminator/D_logit/kernel"], variables["discriminator/D_logit/kernel"])
    assert "discriminator/D_logit/kernel/ExponentialMovingAverage" not in t and "generator/G_logit/u/Adam" not in t
    assert np.array_equal(t["generator/batch_norm/gamma/Adam_1"], adam["generator/batch_norm/gamma"][1])
    assert abs(float(t["beta2_power"]) - 0.9 ** 6) < 1e-7 and abs(float(t["beta2_power_1"]) - 0.9 ** 7) < 1e-7 and float(t["beta1_power"]) == 0.0
    s = ck.from_tf_tensors(t, beta2=0.9)
    assert s["t_g"] == 5 and s["t_d"] == 6
    assert sorted(s["variables"]) == sorted(variables) and all(np.array_equal(s["variables"][n], variables[n]) for n in variables)
    assert all(np.array_equal(s["ema"][n], ema[n]) for n in g_names) and sorted(s["g_adam"]) == sorted(g_names) and sorted(s["d_adam"]) == d_names


def test_step_counts_survive_beta2_power_underflow():
    """beta2_power = 0.9^(t+1) is denormal near t = 830 and exactly 0.0 from t = 1000 in float32: the step counts must not
    come from it there.  Order of preference: the exact counters this package writes, beta2_power while it determines t,
    the '-<step>' suffix of the checkpoint name (BigGAN.py:1267)."""
    v = {"generator/G_logit/kernel": np.zeros((1, 2), np.float32), "discriminator/D_logit/kernel": np.zeros((1, 2), np.float32)}
    adam = {n: (np.zeros((1, 2), np.float32), np.zeros((1, 2), np.float32)) for n in v}
    for t in (2, 500, 979, 1000, 2500, 100000):
        tt = ck.to_tf_tensors(v, ["generator/G_logit/kernel"], ["discriminator/D_logit/kernel"], adam, adam,
                              {"generator/G_logit/kernel": v["generator/G_logit/kernel"]}, t_g=t, t_d=t + 1, beta1=0.0, beta2=0.9)
        s = ck.from_tf_tensors(tt, beta2=0.9, counter=t)
        assert (s["t_g"], s["t_d"]) == (t, t + 1)
        # a checkpoint written by TensorFlow has no exact counters
        tf_only = {k: a for k, a in tt.items() if k not in (ck.STEP_G, ck.STEP_D)}
        s = ck.from_tf_tensors(tf_only, beta2=0.9, counter=t)
        assert s["t_g"] == t, (t, s["t_g"])
        assert s["t_d"] == (t + 1 if t + 1 < 600 else t), (t, s["t_d"])     # past the float32 range: the file-name counter
        assert ck.STEP_G not in s["variables"] and "beta2_power" not in s["variables"]
    assert ck.steps_from_beta_power(0.0, 0.9) is None and ck.steps_from_beta_power(1e-38, 0.9) is None
    assert ck.steps_from_beta_power(0.9 ** 6, 0.9) == 5
    assert ck.counter_from_prefix("/a/b/BigGAN.model-2500") == 2500 and ck.counter_from_prefix("/a/b/weights") is None


def test_checkpoint_state_history_and_pruning(tmp_path):
    """tf.train.Saver(max_to_keep=k) semantics (BigGAN.py:1018): `all_model_checkpoint_paths` keeps the last k prefixes,
    older bundles are deleted; max_to_keep = 0 keeps everything."""
    d = str(tmp_path)
    for step in (10, 20, 30):
        ck.write_bundle(os.path.join(d, "BigGAN.model-%d" % step), {"x": np.array(step, np.int64)})
        ck.write_checkpoint_state(d, "BigGAN.model-%d" % step, max_to_keep=2)
    latest, hist = ck.read_checkpoint_state(d)
    assert latest == "BigGAN.model-30" and hist == ["BigGAN.model-20", "BigGAN.model-30"]
    assert not os.path.exists(os.path.join(d, "BigGAN.model-10.index")) and os.path.exists(os.path.join(d, "BigGAN.model-20.index"))
    assert ck.latest_checkpoint(d) == os.path.join(d, "BigGAN.model-30")
    ck.write_checkpoint_state(d, "BigGAN.model-20", max_to_keep=0)          # re-saving an old step moves it to the end, no duplicates
    assert ck.read_checkpoint_state(d) == ("BigGAN.model-20", ["BigGAN.model-30", "BigGAN.model-20"])
    assert not [f for f in os.listdir(d) if "tmp" in f]

"""Checkpoints in the reference's format and naming.

The reference saves with `self.opt.swapping_saver()` (BigGAN.py:1018, :1255-1284): a TF-1 "V2" checkpoint
(`<prefix>.index` + `<prefix>.data-00000-of-00001`, the TensorBundle format) whose tensors are named after the graph
variables:

    generator/.../kernel, .../bias, .../u, .../pop_mean, ...       the variables
    <var>/Adam, <var>/Adam_1                                       Adam m and v of every trainable variable
    beta1_power, beta2_power / beta1_power_1, beta2_power_1        beta^(t+1) of G's optimizer (created first,
                                                                   BigGAN.py:928-929) / of D's
    <var>/ExponentialMovingAverage                                 MovingAverageOptimizer shadows of G's trainables

and, because the saver is a SWAPPING saver, a generator variable's own name holds its AVERAGED value while
`<var>/ExponentialMovingAverage` holds the raw training value (the two are exchanged on save and on restore).

This module is pure host code: `read_bundle` / `write_bundle` implement the TensorBundle container (a LevelDB-style
table of BundleEntryProto records + one raw data file, CRC-32C everywhere), `to_tf_tensors` / `from_tf_tensors`
map training state to / from that naming, `save` / `load` do both for a Trainer (BigGAN.save / load_checkpoint).  Interop note: written from the published
format (tensorflow/core/util/tensor_bundle, lib/io/table*, lib/hash/crc32c.h); no TensorFlow is available in this
image to cross-check against, so `tests/test_checkpoint.py` pins the format constants, the CRC against RFC 3720
vectors, a hand-assembled index file and the round trip.
"""
import ctypes
import os
import struct

import numpy as np

from . import _lib

TABLE_MAGIC = 0xdb4775248b80fb57           # leveldb table footer
MASK_DELTA = 0xa282ead8                    # crc32c::Mask
RESTART_INTERVAL = 16
BLOCK_SIZE = 256 * 1024                    # BundleWriter's table block size
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8")}
_DT_OF = {np.dtype("float32"): DT_FLOAT, np.dtype("float64"): DT_DOUBLE, np.dtype("int32"): DT_INT32, np.dtype("int64"): DT_INT64}


def crc32c(data, init=0):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    p = a.ctypes.data_as(ctypes.c_void_p) if a.size else None
    return _lib.lib.bg_crc32c(p, int(a.size), int(init)) & 0xFFFFFFFF


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + MASK_DELTA) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------
# varints and the two protos of the bundle (hand-encoded; field numbers of tensor_bundle.proto)
# ---------------------------------------------------------------------------------------------

def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7


def _field(num, wire, payload):
    return _varint((num << 3) | wire) + payload


def encode_header(num_shards=1):
    """BundleHeaderProto{num_shards = 1; endianness = 2 (LITTLE = 0, omitted); version = 3 {producer = 1: 1}}"""
    return _field(1, 0, _varint(num_shards)) + _field(3, 2, _varint(2) + _field(1, 0, _varint(1)))


def encode_entry(dtype, shape, offset, size, crc, shard_id=0):
    """BundleEntryProto{dtype=1, shape=2 {dim=2 {size=1}}, shard_id=3, offset=4, size=5, crc32c=6 (fixed32)}"""
    dims = b"".join(_field(2, 2, (lambda d: _varint(len(d)) + d)(_field(1, 0, _varint(int(s))))) for s in shape)
    out = _field(1, 0, _varint(dtype)) + _field(2, 2, _varint(len(dims)) + dims)
    if shard_id:
        out += _field(3, 0, _varint(shard_id))
    if offset:
        out += _field(4, 0, _varint(offset))
    out += _field(5, 0, _varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


def _decode_fields(buf):
    """-> list of (field number, wire type, value) -- value: int for varint / fixed, bytes for length-delimited."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
        elif wire == 2:
            n, pos = _read_varint(buf, pos)
            v, pos = bytes(buf[pos:pos + n]), pos + n
        elif wire == 5:
            v, pos = struct.unpack_from("<I", buf, pos)[0], pos + 4
        elif wire == 1:
            v, pos = struct.unpack_from("<Q", buf, pos)[0], pos + 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.append((num, wire, v))
    return out


def decode_entry(buf):
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "slices": False}
    for num, _, v in _decode_fields(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            for n2, _, d in _decode_fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, s in _decode_fields(d):
                        if n3 == 1:
                            size = s
                    e["shape"].append(size)
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = v
        elif num == 7:
            e["slices"] = True
    return e


# ---------------------------------------------------------------------------------------------
# the table (tensorflow/core/lib/io/table: blocks with prefix-compressed keys + restart points)
# ---------------------------------------------------------------------------------------------

def _build_block(items):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % RESTART_INTERVAL == 0:
            restarts.append(len(out))
        else:
            n = min(len(k), len(last))
            while shared < n and k[shared] == last[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _parse_block(buf):
    n_restarts = struct.unpack_from("<I", buf, len(buf) - 4)[0]
    end = len(buf) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(buf, pos)
        unshared, pos = _read_varint(buf, pos)
        vlen, pos = _read_varint(buf, pos)
        key = key[:shared] + bytes(buf[pos:pos + unshared])
        pos += unshared
        out.append((key, bytes(buf[pos:pos + vlen])))
        pos += vlen
    return out


def _block_with_trailer(block):
    trailer_type = b"\x00"                                  # kNoCompression
    return block + trailer_type + struct.pack("<I", mask_crc(crc32c(block + trailer_type)))


def _handle(offset, size):
    return _varint(offset) + _varint(size)


def build_table(items):
    """items: sorted list of (key bytes, value bytes) -> the bytes of a table file."""
    out, index = bytearray(), []
    cur, cur_bytes = [], 0

    def flush():
        nonlocal cur, cur_bytes
        if not cur:
            return
        block = _build_block(cur)
        index.append((cur[-1][0], _handle(len(out), len(block))))      # index key: any key >= the block's last key
        out.extend(_block_with_trailer(block))
        cur, cur_bytes = [], 0

    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 8
        if cur_bytes >= BLOCK_SIZE:
            flush()
    flush()
    meta = _build_block([])
    meta_handle = _handle(len(out), len(meta))
    out.extend(_block_with_trailer(meta))
    idx = _build_block(index)
    idx_handle = _handle(len(out), len(idx))
    out.extend(_block_with_trailer(idx))
    footer = meta_handle + idx_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    return bytes(out)


def parse_table(buf, verify=True):
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("not a TensorFlow checkpoint index (bad table magic)")
    footer = buf[len(buf) - 48:]
    _, p = _read_varint(footer, 0)
    _, p = _read_varint(footer, p)
    ioff, p = _read_varint(footer, p)
    isize, p = _read_varint(footer, p)

    def block(off, size):
        raw = buf[off:off + size]
        ctype = buf[off + size]
        if ctype != 0:
            raise ValueError("compressed table blocks (type %d) are not supported" % ctype)
        if verify:
            stored = struct.unpack_from("<I", buf, off + size + 1)[0]
            if stored != mask_crc(crc32c(bytes(raw) + b"\x00")):
                raise ValueError("checkpoint index: block checksum mismatch at offset %d" % off)
        return raw

    items = []
    for _, h in _parse_block(block(ioff, isize)):
        off, p = _read_varint(h, 0)
        size, _ = _read_varint(h, p)
        items += _parse_block(block(off, size))
    return items


# ---------------------------------------------------------------------------------------------
# bundle = index table + data shard(s)
# ---------------------------------------------------------------------------------------------

def write_bundle(prefix, tensors):
    """tensors: {name: ndarray (float32 / float64 / int32 / int64)} -> <prefix>.index, <prefix>.data-00000-of-00001."""
    names = sorted(tensors, key=lambda s: s.encode("utf-8"))
    items = [(b"", encode_header(1))]
    offset = 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.asarray(tensors[n])                       # (np.ascontiguousarray would turn a scalar into shape [1])
            if not a.flags.c_contiguous:
                a = a.copy(order="C")
            if a.dtype not in _DT_OF:
                raise ValueError("unsupported dtype %s for %s" % (a.dtype, n))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            items.append((n.encode("utf-8"), encode_entry(_DT_OF[a.dtype], a.shape, offset, len(raw), mask_crc(crc32c(raw)))))
            offset += len(raw)
    with open(prefix + ".index", "wb") as f:
        f.write(build_table(items))


def read_bundle(prefix, verify=True, names=None):
    """-> {name: ndarray}.  Verifies block and tensor checksums unless verify=False."""
    with open(prefix + ".index", "rb") as f:
        items = parse_table(f.read(), verify)
    if not items or items[0][0] != b"":
        raise ValueError("checkpoint index has no header entry")
    num_shards = 1
    for num, _, v in _decode_fields(items[0][1]):
        if num == 1:
            num_shards = v
        if num == 2 and v != 0:
            raise ValueError("big-endian checkpoints are not supported")
    shards = {}
    out = {}
    for k, v in items[1:]:
        name = k.decode("utf-8")
        if names is not None and name not in names:
            continue
        e = decode_entry(v)
        if e["slices"]:
            raise ValueError("partitioned variable %s is not supported" % name)
        if e["dtype"] not in _DTYPES:
            raise ValueError("tensor %s has unsupported dtype enum %d" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, num_shards), dtype=np.uint8, mode="r")
        raw = np.asarray(shards[sid][e["offset"]:e["offset"] + e["size"]])
        if verify and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise ValueError("checkpoint data: checksum mismatch for %s" % name)
        out[name] = raw.view(_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def read_checkpoint_state(directory):
    """-> (model_checkpoint_path, [all_model_checkpoint_paths]) of the text `checkpoint` file, or (None, [])."""
    path = os.path.join(directory, "checkpoint")
    latest, history = None, []
    if os.path.exists(path):
        for line in open(path):
            key, _, val = line.partition(":")
            val = val.strip().strip('"')
            if key.strip() == "model_checkpoint_path":
                latest = val
            elif key.strip() == "all_model_checkpoint_paths":
                history.append(val)
    return latest, history


def write_checkpoint_state(directory, prefix_basename, max_to_keep=0):
    """The text `checkpoint` file tf.train.get_checkpoint_state reads (BigGAN.py:1279).  Like tf.train.Saver it keeps the
    history of saved prefixes in `all_model_checkpoint_paths`; with max_to_keep > 0 (the reference passes
    --keep_checkpoints, BigGAN.py:1018) the oldest prefixes beyond that count are dropped from the list and their files
    deleted.  The state file is replaced atomically."""
    _, history = read_checkpoint_state(directory)
    history = [h for h in history if h != prefix_basename] + [prefix_basename]
    if max_to_keep and max_to_keep > 0:
        for old in history[:-max_to_keep]:
            base = old if os.path.isabs(old) else os.path.join(directory, old)
            for fn in _bundle_files(base):
                if os.path.exists(fn):
                    os.remove(fn)
        history = history[-max_to_keep:]
    tmp = os.path.join(directory, "checkpoint.tmp.%d" % os.getpid())
    with open(tmp, "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % prefix_basename)
        for h in history:
            f.write('all_model_checkpoint_paths: "%s"\n' % h)
    os.replace(tmp, os.path.join(directory, "checkpoint"))


def _bundle_files(prefix):
    return [prefix + ".index", prefix + ".data-00000-of-00001"]


def latest_checkpoint(directory):
    path = os.path.join(directory, "checkpoint")
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("model_checkpoint_path:"):
            p = line.split(":", 1)[1].strip().strip('"')
            return p if os.path.isabs(p) else os.path.join(directory, p)
    return None


# ---------------------------------------------------------------------------------------------
# naming: Trainer state <-> the tensors of the reference's swapping saver
# ---------------------------------------------------------------------------------------------

def to_tf_tensors(variables, g_trainable, d_trainable, g_adam, d_adam, ema, t_g, t_d, beta1, beta2):
    """variables: {name: array} (all of them: weights, u, BN statistics); g_adam / d_adam: {name: (m, v)};
    ema: {name: shadow} for G's trainables; t_g / t_d: optimizer steps taken.  Returns the checkpoint's tensors."""
    out = {}
    for n, a in variables.items():
        out[n] = np.asarray(a, dtype=np.float32)
    for n in g_trainable:                                   # swapped: the variable's name carries the average
        out[n] = np.asarray(ema[n], dtype=np.float32)
        out[n + "/ExponentialMovingAverage"] = np.asarray(variables[n], dtype=np.float32)
    for names, adam in ((g_trainable, g_adam), (d_trainable, d_adam)):
        for n in names:
            out[n + "/Adam"] = np.asarray(adam[n][0], dtype=np.float32)
            out[n + "/Adam_1"] = np.asarray(adam[n][1], dtype=np.float32)
    out["beta1_power"] = np.array(beta1 ** (t_g + 1), dtype=np.float32)        # G's optimizer is built first
    out["beta2_power"] = np.array(beta2 ** (t_g + 1), dtype=np.float32)
    out["beta1_power_1"] = np.array(beta1 ** (t_d + 1), dtype=np.float32)
    out["beta2_power_1"] = np.array(beta2 ** (t_d + 1), dtype=np.float32)
    # the float32 beta2_power underflows (denormal near t = 830, exactly 0 from t = 1000 at beta2 = 0.9), so the exact
    # optimizer step counts travel as two extra tensors; TF's restore only reads the names its graph asks for.
    out[STEP_G] = np.array(t_g, dtype=np.int64)
    out[STEP_D] = np.array(t_d, dtype=np.int64)
    return out


STEP_G = "biggan_b200/adam_step_generator"
STEP_D = "biggan_b200/adam_step_discriminator"


def counter_from_prefix(prefix):
    """BigGAN.load_checkpoint (BigGAN.py:1267): the iteration counter is the '-<step>' suffix of the checkpoint name."""
    try:
        return int(os.path.basename(prefix).split('-')[-1])
    except ValueError:
        return None


def steps_from_beta_power(p, beta2):
    """t from beta2_power = beta2^(t+1), or None where the float32 value no longer determines t (outside (0, 1),
    or denormal / close to it: below 1e-30 one ulp spans several steps)."""
    p = float(p)
    if not (1e-30 < p < 1.0) or not (0.0 < beta2 < 1.0):
        return None
    return max(int(round(np.log(p) / np.log(beta2))) - 1, 0)


def from_tf_tensors(tensors, beta2, counter=None):
    """Inverse of to_tf_tensors -> dict(variables, g_adam, d_adam, ema, t_g, t_d).  Optimizer step counts, in order of
    preference: the exact counters this package writes; beta2_power while it still determines t (cross-checked against
    nothing else -- a checkpoint written by TensorFlow has only this); the iteration counter of the file name
    (BigGAN.py:1267; one D and one G step per iteration at n_critic = 1), which is what the reference resumes from."""
    skip_suffix = ("/Adam", "/Adam_1", "/ExponentialMovingAverage")
    special = {"beta1_power", "beta2_power", "beta1_power_1", "beta2_power_1", STEP_G, STEP_D}
    variables = {n: a for n, a in tensors.items() if not n.endswith(skip_suffix) and n not in special}
    ema, g_adam, d_adam = {}, {}, {}
    for n in list(variables):
        if n + "/ExponentialMovingAverage" in tensors:     # swap back
            ema[n] = variables[n]
            variables[n] = tensors[n + "/ExponentialMovingAverage"]
        if n + "/Adam" in tensors:
            (g_adam if "generator" in n else d_adam)[n] = (tensors[n + "/Adam"], tensors[n + "/Adam_1"])

    def steps(own, key):
        if own in tensors:
            return int(tensors[own])
        t = steps_from_beta_power(tensors[key], beta2) if key in tensors else None
        if t is not None:
            return t
        return int(counter) if counter is not None else 0

    return {"variables": variables, "g_adam": g_adam, "d_adam": d_adam, "ema": ema,
            "t_g": steps(STEP_G, "beta2_power"), "t_d": steps(STEP_D, "beta2_power_1")}


def _arena_slices(trainer, net, flat):
    ar = trainer.store.arenas[net]
    host = flat.detach().cpu().numpy()
    return {v.name: host[o:o + v.data.numel()].reshape(tuple(v.data.shape)).copy() for v, o in zip(ar["vars"], ar["offsets"])}


def save(trainer, directory, step, model_name="BigGAN", max_to_keep=0):
    """BigGAN.save (BigGAN.py:1255-1261): <directory>/<model_name>.model-<step>.{index,data-00000-of-00001} + `checkpoint`
    (history pruned to max_to_keep prefixes when > 0, the reference's --keep_checkpoints).

    Data parallel: COLLECTIVE -- every rank must call it (with a sharded optimizer the Adam / EMA slices are all-gathered).
    Replicas are identical, so rank 0 alone writes; files are written under a temporary name and renamed into place, and
    all ranks leave through a barrier so that no rank reads a half-written checkpoint."""
    import torch
    if not trainer._built:
        trainer.build()
    torch.cuda.synchronize()
    cfg = trainer.cfg
    full = {(net, key): trainer.full_opt(net, key) for net in ("generator", "discriminator") for key in ("m", "v")}   # collective
    full_ema = trainer.full_ema()                                                                                       # collective
    base = "%s.model-%d" % (model_name, step)
    if trainer.rank == 0:
        variables = trainer.store.export()
        adam = {}
        for net in ("generator", "discriminator"):
            m, v = _arena_slices(trainer, net, full[(net, "m")]), _arena_slices(trainer, net, full[(net, "v")])
            adam[net] = {n: (m[n], v[n]) for n in m}
        ema = _arena_slices(trainer, "generator", full_ema)
        t_g = int(round(float(trainer.opt["generator"]["state"][0].cpu())))
        t_d = int(round(float(trainer.opt["discriminator"]["state"][0].cpu())))
        tensors = to_tf_tensors(variables, list(adam["generator"]), list(adam["discriminator"]), adam["generator"],
                                adam["discriminator"], ema, t_g, t_d, cfg.beta1, cfg.beta2)
        os.makedirs(directory, exist_ok=True)
        tmp = os.path.join(directory, ".tmp-%d-%s" % (os.getpid(), base))
        write_bundle(tmp, tensors)
        for src, dst in zip(_bundle_files(tmp), _bundle_files(os.path.join(directory, base))):
            os.replace(src, dst)
        write_checkpoint_state(directory, base, max_to_keep=max_to_keep)
    if trainer.world > 1:
        import torch.distributed as dist
        dist.barrier(group=trainer.pg)
    return os.path.join(directory, base)


def load(trainer, prefix):
    """BigGAN.load_checkpoint (BigGAN.py:1264-1270): restore weights, state, Adam slots, EMA shadows and step counts.
    Returns the iteration counter to resume from: the '-<step>' suffix of the checkpoint name, as the reference does."""
    import torch
    if not trainer._built:
        trainer.build()
    cfg = trainer.cfg
    counter = counter_from_prefix(prefix)
    st = from_tf_tensors(read_bundle(prefix), cfg.beta2, counter=counter)
    missing = sorted(set(trainer.store.vars) - set(st["variables"]))
    if missing:
        raise KeyError("checkpoint %s lacks %d variables, e.g. %s" % (prefix, len(missing), missing[:3]))
    trainer.store.load({n: st["variables"][n] for n in trainer.store.vars})
    for net, adam, t in (("generator", st["g_adam"], st["t_g"]), ("discriminator", st["d_adam"], st["t_d"])):
        ar, o = trainer.store.arenas[net], trainer.opt[net]
        for v, off in zip(ar["vars"], ar["offsets"]):
            n = v.data.numel()
            if v.name in adam:
                o["m"][off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(adam[v.name][0], np.float32)).reshape(-1))
                o["v"][off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(adam[v.name][1], np.float32)).reshape(-1))
            if net == "generator" and v.name in st["ema"]:
                o["ema"][off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(st["ema"][v.name], np.float32)).reshape(-1))
        lr = cfg.g_lr if net == "generator" else cfg.d_lr
        lr_t = lr * np.sqrt(1 - cfg.beta2 ** t) / (1 - cfg.beta1 ** t) if t > 0 else 0.0
        o["state"].copy_(torch.tensor([float(t), float(lr_t)]))
    trainer.t_g, trainer.t_d = st["t_g"], st["t_d"]
    return counter if counter is not None else st["t_g"]

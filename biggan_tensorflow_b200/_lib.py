"""ctypes binding of libbiggan_b200.so (the C ABI declared in include/biggan_b200.h).

The argument types are parsed from the header itself, so the Python side can never drift from the
ABI.  There is NO CPU fallback: if the shared library is missing this module raises at import.
"""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
_HEADER = os.path.join(os.path.dirname(_PKG), "include", "biggan_b200.h")
_SO = os.path.join(_PKG, "libbiggan_b200.so")

BG_F32, BG_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
PAD_ZERO, PAD_REFLECT = 0, 1


class ConvDesc(ctypes.Structure):
    """mirror of bg_conv_desc"""
    _fields_ = [(n, ctypes.c_int) for n in
                ("B", "Hi", "Wi", "Ci", "Ho", "Wo", "Co", "kh", "kw", "stride",
                 "pad_t", "pad_l", "pad_b", "pad_r", "pad_mode", "transposed")]


_CTYPES = {
    "int": ctypes.c_int, "long": ctypes.c_long, "size_t": ctypes.c_size_t,
    "float": ctypes.c_float, "double": ctypes.c_double,
}


def parse_header(path=_HEADER):
    """-> {name: (restype, [argtypes])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"#[^\n]*", "", src)
    src = re.sub(r"typedef\s+(enum|struct)\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(const\s+char\s*\*|size_t|int)\s+(bg_\w+)\s*\(([^)]*)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if ret.startswith("const"):
            restype = ctypes.c_char_p
        elif ret == "size_t":
            restype = ctypes.c_size_t
        else:
            restype = ctypes.c_int
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    toks = [t for t in a.split() if t != "const"]
                    argtypes.append(_CTYPES[toks[0]])
        decls[name] = (restype, argtypes)
    return decls


DECLS = parse_header()

if not os.path.exists(_SO):
    raise RuntimeError(
        "libbiggan_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C biggan_tensorflow_b200/csrc`.  There is no CPU fallback." % _SO)

lib = ctypes.CDLL(_SO)
for _name, (_res, _args) in DECLS.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch
    _fn.restype = _res
    _fn.argtypes = _args

launch_count = 0   # kernels-launching C calls issued (bench.py reports it as gpu_launches)


class BgError(RuntimeError):
    pass


def _conv(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    if isinstance(a, ctypes.Structure):
        return ctypes.addressof(a)
    return a


_prof = None       # when a list: per-call CUDA-event timing records (name, start, end, flops, tag)


def profile_begin():
    global _prof
    _prof = []


def profile_end():
    """-> {name: {"calls", "ms", "flops"}} aggregated over the recorded calls (synchronises)."""
    global _prof
    import torch
    recs, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, flops in recs:
        d = out.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0})
        d["calls"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += flops
    return out


class timed:
    """with timed("nccl_allreduce_grads"): ...  -- CUDA-event record of a non-C-ABI region (collectives) while profiling."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if _prof is not None:
            import torch
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _prof is not None:
            self.e1.record()
            _prof.append((self.tag, self.e0, self.e1, 0.0))
        return False


def call(name, *args, flops=0.0, tag=None):
    """Invoke a status-returning entry point; raise with bg_last_error() on failure."""
    global launch_count
    fn = getattr(lib, name)
    cargs = [_conv(a) for a in args]
    if _prof is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*cargs)
        e1.record()
        _prof.append((tag or name, e0, e1, flops))
    else:
        rc = fn(*cargs)
    launch_count += 1
    if rc != 0:
        msg = lib.bg_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise ValueError("%s: %s" % (name, msg))
        raise BgError("%s failed (%d): %s" % (name, rc, msg))


def query(name, *args):
    """Invoke a value-returning entry point (workspace sizes etc.)."""
    return getattr(lib, name)(*[_conv(a) for a in args])

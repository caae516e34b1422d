"""TEST INFRASTRUCTURE ONLY: point the Python layer of biggan_tensorflow_b200 at the host library (host_lib.cpp) and at
CPU tensors.  `activate(so, setattr_fn)` is used with pytest's monkeypatch.setattr inside fixtures (undone after the test)
and with the builtin setattr inside spawned worker processes (the gloo data-parallel test).  The product package has no
switch for this: every name patched here is patched from the outside."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(so):
    subprocess.run(["g++", "-std=c++20", "-O2", "-x", "c++", "-shared", "-fPIC", os.path.join(HERE, "host_lib.cpp"), "-o", so], check=True)
    return so


def activate(so, set_attr=setattr):
    import torch
    from biggan_tensorflow_b200 import DiffAugment_tf, _lib, ops, sn_group, step, tensor, tf_shim, utils
    lib = ctypes.CDLL(so)
    for name, (res, args) in _lib.DECLS.items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    set_attr(_lib, "lib", lib)
    for mod in (tensor, tf_shim, step, sn_group, DiffAugment_tf):
        set_attr(mod, "device", lambda: torch.device("cpu"))
    for mod in (tensor, ops, utils, tf_shim, step, sn_group, DiffAugment_tf):
        set_attr(mod, "stream", lambda: None)
    set_attr(tensor, "_ws_cache", {})
    set_attr(torch.cuda, "current_device", lambda: 0)
    set_attr(torch.cuda, "synchronize", lambda *a, **k: None)
    DiffAugment_tf._queue.clear()
    return lib

"""Diagnostic: error of every checked quantity of every golden op case on the CUDA path (no assertions).
   python tools/op_errors.py [f32|bf16] [threshold]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_zz_reference_golden_gpu as T                            # noqa: E402
C = T.C
dtype = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float32
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
log = []


def check(gold, key, value, tol, what=None, l2=False):
    v = np.asarray(value, dtype=np.float64)
    if key in gold.files:
        ref = np.asarray(gold[key], dtype=np.float64)
        v = v.reshape(ref.shape)
    else:
        stride = int(gold[key + "#stat"][1])
        ref = np.asarray(gold[key + "#sub"], dtype=np.float64)
        v = v.ravel()[::stride]
    scale = max(float(np.abs(ref).max()), 1e-6)
    e_max = float(np.abs(v - ref).max()) / scale
    e_l2 = float(np.linalg.norm(v - ref)) / max(float(np.linalg.norm(ref)), 1e-30)
    log.append((key, e_max, e_l2, tol))


C.check = check
for cases, fname in ((C.OP_CASES, "ref_ops.npz"), (C.NEXT_OP_CASES, "ref_ops_next.npz")):
    for case in cases:
        if dtype == torch.bfloat16 and case["name"] in T.FP32_ONLY:
            continue
        try:
            T._cuda_op_case(case, dtype, fname)
        except Exception as e:                                      # noqa: BLE001
            print("!!", case["name"], repr(e)[:200])
for key, e_max, e_l2, tol in log:
    if e_max >= thr:
        print("%-70s max-abs/max|ref| %.2e   rel-L2 %.2e   (tol %.1e)" % (key, e_max, e_l2, tol))

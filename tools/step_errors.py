"""Diagnostic: per-variable relative L2 error of the bf16 (or fp32) step gradients vs the oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_step_gpu import _run
dtype = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
r = _run(dtype, "")
print("d_loss %.6f ref %.6f   g_loss %.6f ref %.6f" % (r["d_loss"], r["rd_loss"], r["g_loss"], r["rg_loss"]))
for which, ref, got in (("D", r["rd_grads"], r["d_grads"]), ("G", r["rg_grads"], r["g_grads"])):
    for name, g in ref.items():
        g = g.numpy()
        e = np.linalg.norm(got[name] - g) / max(np.linalg.norm(g), 1e-30)
        flag = " <<<<" if e > 4e-2 and not name.endswith("f_conv/bias") else ""
        print("%s %-70s relL2 %.3e |ref| %.3e%s" % (which, name, e, np.linalg.norm(g), flag))

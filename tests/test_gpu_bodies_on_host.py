"""The bodies of the GPU tests of the training step (tests/test_step_gpu.py, tests/test_golden.py: one full D + G
iteration with Adam / EMA applied, with and without DiffAugment, against the oracle and against the oracle-generated
golden step; fp32 and bf16 storage mode) executed on the test-only host implementation of the C ABI.  Purpose: a CPU-side regression guard for the
Python layer of the canonical path, and a cross-check of the host library itself against tests that are known to pass
on hardware."""
import pytest
import torch

from test_next_ops_glue_host import host_env                    # noqa: F401
import test_golden as TG
import test_step_gpu as S


@pytest.fixture()
def host_all(host_so, monkeypatch):
    from biggan_tensorflow_b200 import tensor, tf_shim
    host_env.activate(host_so, monkeypatch.setattr)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    yield
    tf_shim.set_store(tf_shim.VariableStore())
    tensor.clear_tape()


@pytest.mark.parametrize("mode,da_policy", [("fp32", ""), ("bf16", "color,translation,cutout"), ("bf16-tcglue", ""),
                                            ("bf16-tcglue", "color,translation,cutout")])
def test_train_step_matches_oracle_on_the_host(host_all, monkeypatch, da_policy, mode):
    """bf16-tcglue: the host library reports a tensor-core path (BG_HOST_TC=1), so the Python layer takes its tensor-core-mode
    glue -- 16-channel image layout, im2col of the RGB convs, per-tap transposed weight copies, zero-padded narrow
    projections, the fused attention entry points -- over the same loop stand-ins."""
    if mode == "bf16-tcglue":
        from biggan_tensorflow_b200 import ops
        monkeypatch.setenv("BG_HOST_TC", "1")
        ops.set_compute_dtype(torch.bfloat16)
        assert ops.rgb_pad() == 16
    S.test_train_step_matches_oracle(torch.float32 if mode == "fp32" else torch.bfloat16, da_policy)


def test_parameters_after_update_on_the_host(host_all):
    S.test_parameters_after_update_fp32()


def test_step_matches_oracle_golden_vectors_on_the_host(host_all):
    TG.test_cuda_step_matches_golden_vectors()


@pytest.mark.parametrize("case", [(2, 8, 32, "reflect", True), (2, 33, 16, "reflect", False), (2, 30, 64, "zero", True), (3, 2, 16, "reflect", False)])
def test_rgb_output_conv_tap_stacked_on_the_host(host_all, monkeypatch, case):
    """tests/test_ops_gpu.py::test_conv_rgb_out_tap_stacked on the host library (BG_HOST_TC=1): the Python glue of the tap-stacked
    G_logit path and the arithmetic of its four entry points (loop stand-ins in host_lib.cpp) against the oracle."""
    from biggan_tensorflow_b200 import ops, _lib
    import test_ops_gpu as O
    monkeypatch.setenv("BG_HOST_TC", "1")
    ops.set_compute_dtype(torch.bfloat16)
    seen = []
    real_call = _lib.call
    monkeypatch.setattr(ops, "call", lambda name, *a, **k: (seen.append(name), real_call(name, *a, **k))[1])
    O.rgb_out_body(case)
    for name in ("bg_rgb_stack_weights", "bg_rgb_tapsum_fwd", "bg_rgb_tapsum_bwd", "bg_rgb_unstack_wgrad"):
        assert name in seen, name
    seen.clear()
    monkeypatch.setenv("BG_RGB_TAPSUM", "0")               # the plain 3x3 path gives the same answer
    O.rgb_out_body(case)
    assert "bg_rgb_tapsum_fwd" not in seen


"""The bodies of the GPU tests of the training step (tests/test_step_gpu.py, tests/test_golden.py: one full D + G
iteration with Adam / EMA applied, with and without DiffAugment, against the oracle and against the oracle-generated
golden step; fp32 and bf16 storage mode) executed on the test-only host implementation of the C ABI.  Purpose: a CPU-side regression guard for the
Python layer of the canonical path, and a cross-check of the host library itself against tests that are known to pass
on hardware."""
import pytest
import torch

from test_next_ops_glue_host import host_env, host_so           # noqa: F401  (fixture)
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


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("da_policy", ["", "color,translation,cutout"])
def test_train_step_matches_oracle_on_the_host(host_all, da_policy, dtype):
    S.test_train_step_matches_oracle(dtype, da_policy)


def test_parameters_after_update_on_the_host(host_all):
    S.test_parameters_after_update_fp32()


def test_step_matches_oracle_golden_vectors_on_the_host(host_all):
    TG.test_cuda_step_matches_golden_vectors()

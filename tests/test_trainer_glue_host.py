"""`step.Trainer` (arenas, grouped spectral norm, the D / G update phases, regulariser gradients, bn_in_d's two D passes,
virtual batches, EMA sampling, checkpoint round trip) executed WITHOUT a GPU in fp32: `_lib.lib` points at the host
library of tests/host_emul/host_lib.cpp (loop stand-ins written from include/biggan_b200.h + the real source of the 8(f)
kernels), device / stream helpers at the CPU.  The bodies are the SAME functions the GPU tests of
tests/test_zz_reference_golden_gpu.py run, so this also exercises that test code; expected values are the vectors produced
by the reference's own generator() / discriminator().  It checks the Python side of the training step, not the kernels."""
import os

import pytest
import torch

from test_next_ops_glue_host import host_env                     # noqa: F401
import test_zz_reference_golden_gpu as G


@pytest.fixture()
def host_trainer(host_so, monkeypatch):
    from biggan_tensorflow_b200 import tensor, tf_shim
    host_env.activate(host_so, monkeypatch.setattr)
    yield lambda a: torch.from_numpy(a)
    tf_shim.set_store(tf_shim.VariableStore())
    tensor.clear_tape()


@pytest.mark.parametrize("name", ["net64", "net128da", "net512", "net64cli", "net64variants"])     # net256: tests/test_model_wiring_host.py
def test_trainer_phases_match_reference_networks_on_the_host(host_trainer, name):
    """One D phase and one G phase (losses, every gradient norm, selected gradients, all state after each phase, variable
    names and shapes) for the canonical networks at 64, 128 (+ DiffAugment), 256 and 512 pixels (BASELINE configs 1-4 at narrow
    width), the fork's CLI defaults (prelu, ortho_cosine, flood, DiffAugment) and the layer variants (lrelu, resize_conv,
    conv3, resize_conv1, bn_in_d, ortho)."""
    G._cuda_net(name, dev=host_trainer, split=(name == "net64"))     # the float64 split check once is enough on the host


def test_virtual_batches_and_ema_sampling_on_the_host(host_trainer):
    G._virtual_batches_and_sampling(host_trainer)


def test_checkpoint_round_trip_on_the_host(host_trainer, tmp_path):
    G._checkpoint_round_trip(tmp_path, host_trainer)

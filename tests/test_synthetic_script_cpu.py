"""CPU tests of the synthetic-function benchmark mirror (synthetic_test_script.py; BASELINE configs[0] is a point of its learning
curves): target functions against fixtures from the reference's own definitions (tests/golden/make_golden.py synth), model
constructors lower to the fused operator, nothing runs at import time.  The training loop needs a GPU
(tests/test_model_gpu.py::test_synthetic_benchmark_point)."""
import os

import numpy as np
import pytest
import torch

import synthetic_test_script as sts
from rpgp.lazy import RPAdditiveLazyTensor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_target_functions_match_the_reference_definitions():
    g = np.load(os.path.join(GOLD, "synthetic_targets.npz"))
    x = torch.from_numpy(g["x"])
    assert set(sts.TARGETS) == set(g.files) - {"x"}
    for name, fn in sts.TARGETS.items():
        np.testing.assert_allclose(fn(x.clone()).numpy(), g[name], rtol=1e-6, atol=1e-7, err_msg=name)


@pytest.mark.parametrize("name,kwargs,J,K,base", [("gam", {}, 6, 1, 0), ("dpa_ard", {"J": 6}, 6, 1, 0), ("poly_rp", {"J": 4, "k": 1}, 4, 1, 0),
                                                  ("rp", {}, 6, 1, 1), ("bl", {}, 1, 6, 1)])
def test_model_constructors_lower_to_the_fused_operator(name, kwargs, J, K, base):
    torch.manual_seed(0)
    np.random.seed(0)
    x = torch.rand(12, 6) * 4 - 2
    model = sts.MODELS[name](x, sts.additive(x), **kwargs)
    op = model.covar_module(x).evaluate_kernel()
    assert isinstance(op, RPAdditiveLazyTensor) and (op.J, op.K, op.base) == (J, K, base) and op.symmetric
    assert tuple(op.Z1.shape) == (12, J * K)


def test_sizes_and_cli_defaults():
    assert sts.SIZES[0] == 10 and sts.SIZES[-1] == 10240 and all(b == 2 * a for a, b in zip(sts.SIZES, sts.SIZES[1:]))
    assert sts.device.startswith("cuda")

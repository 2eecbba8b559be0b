"""CPU tests of the base-kernel plumbing: oracle functions and their slopes, the torch restatement used for
diagonals, the kernel-type lookup of training_routines, and the rules for combining operators."""
import numpy as np
import pytest
import torch

import training_routines as tr
from gp_models.kernels import InverseMQKernel as MirrorIMQ, KeOpsInverseMQKernel, postprocess_inverse_mq
from oracle import rpgp_oracle as orc
from rpgp import lazy
from rpgp.gp import kernels as gk


@pytest.mark.parametrize("base", [0, 1, 2, 3])
def test_oracle_base_functions_and_slopes(base):
    sq = np.linspace(0.0, 30.0, 301)[1:]
    h = 1e-6
    fd = (orc.base_f(base, sq + h) - orc.base_f(base, sq - h)) / (2 * h)
    np.testing.assert_allclose(orc.base_df(base, sq), fd, rtol=1e-6, atol=1e-10)
    assert abs(orc.base_f(base, np.array([0.0]))[0] - 1.0) < 1e-15            # k(x, x) = 1 for all four
    assert abs(orc.base_df(base, np.array([0.0]))[0] - (-1.5 if base == 1 else -0.5)) < 1e-12      # finite slope at the origin
    t = torch.from_numpy(sq)
    np.testing.assert_allclose(lazy.base_function(base, t).numpy(), orc.base_f(base, sq), rtol=1e-12)


def test_oracle_dense_kernel_known_values():
    # one group, one coordinate, distance 2: RBF e^-2, Matern (1 + 2 sqrt3) e^(-2 sqrt3), IMQ 5^-1/2, cosine cos 2
    Z1, Z2 = np.array([[0.0]]), np.array([[2.0]])
    want = [np.exp(-2.0), (1 + 2 * np.sqrt(3)) * np.exp(-2 * np.sqrt(3)), 5 ** -0.5, np.cos(2.0)]
    for base in range(4):
        assert abs(orc.additive_rbf_dense(Z1, Z2, [1.0], 1, 1, base=base)[0, 0] - want[base]) < 1e-15
    assert abs(float(postprocess_inverse_mq(torch.tensor(4.0))) - want[2]) < 1e-7      # imq_kernel.py:8-9


def test_kernel_type_lookup_matches_the_reference_table():
    assert tr._map_to_kernel(False, "RBF", False) == (gk.RBFKernel, {})
    assert tr._map_to_kernel(False, "Matern", True) == (gk.MaternKernel, {"nu": 1.5})
    assert tr._map_to_kernel(False, "InverseMQ", False) == (gk.InverseMQKernel, {})
    assert MirrorIMQ is gk.InverseMQKernel and KeOpsInverseMQKernel is gk.InverseMQKernel
    assert gk.keops.RBFKernel is gk.RBFKernel and gk.keops.MaternKernel is gk.MaternKernel
    assert tr._map_to_kernel(False, "Cosine", False) == (gk.CosineKernel, {})       # (dense-only in the reference, :76-81)
    with pytest.raises(ValueError):
        tr._map_to_kernel(False, "Periodic", False)
    k = tr._map_to_kernel(True, "Matern", False, active_dims=[0])
    assert isinstance(k, gk.MaternKernel) and k.nu == 1.5 and tuple(k.raw_lengthscale.shape) == (1, 1)
    with pytest.raises(NotImplementedError):
        gk.MaternKernel(nu=2.5)


def test_operators_of_different_base_kernels_do_not_merge():
    Z = torch.zeros(4, 2)
    a = lazy.RPAdditiveLazyTensor(Z, None, torch.ones(2), 2, 1, base=1)
    b = lazy.RPAdditiveLazyTensor(Z, None, torch.ones(2), 2, 1, base=1)
    s = a + b
    assert s.base == 1 and s.J == 4 and s._transpose_nonbatch().base == 1 and s.scale(2.0).base == 1
    with pytest.raises(NotImplementedError):
        lazy.RPAdditiveLazyTensor.sum([a, lazy.RPAdditiveLazyTensor(Z, None, torch.ones(2), 2, 1, base=0)])
    one = lazy.RPAdditiveLazyTensor(Z[:, :1], None, torch.ones(1), 1, 1, base=2)
    with pytest.raises(NotImplementedError):          # a product of non-RBF kernels is not a kernel of the summed distance
        lazy.RPAdditiveLazyTensor.product([one, one])
    op = gk.InverseMQKernel().forward(Z, Z)
    assert op.base == 2


def test_cosine_kernel_is_gpytorchs_cosine_kernel():
    """gpytorch.kernels.CosineKernel: cos(pi |a - b| / period_length), one positive `raw_period_length` of shape (1, 1), no
    lengthscale; the reference's additive factory initialises period_length = 1 where the others get lengthscale = 1
    (training_routines.py:150-153).  The class lowers to base kernel 3 with pi / p folded into the coordinates."""
    k = gk.CosineKernel()
    assert tuple(k.raw_period_length.shape) == (1, 1) and k.lengthscale is None and not k.has_lengthscale
    with pytest.raises(RuntimeError):
        k.lengthscale = 1.0
    k.initialize(period_length=torch.tensor([2.5]))
    np.testing.assert_allclose(float(k.period_length), 2.5, rtol=1e-6)
    g = torch.Generator().manual_seed(0)
    x1, x2 = torch.randn(7, 3, generator=g, dtype=torch.float64), torch.randn(5, 3, generator=g, dtype=torch.float64)
    k = k.double()
    k.initialize(period_length=torch.tensor([2.5], dtype=torch.float64))
    op = k.forward(x1, x2)
    assert op.base == 3 and (op.J, op.K) == (1, 3)
    dense = orc.additive_rbf_dense(op.Z1.detach().numpy(), op.Z2.detach().numpy(), [1.0], 1, 3, base=3)
    want = np.cos(np.pi * torch.cdist(x1, x2).numpy() / 2.5)
    np.testing.assert_allclose(dense, want, rtol=1e-10, atol=1e-12)
    per_dim = k.forward(x1, x1, last_dim_is_batch=True)           # AdditiveStructureKernel's way: one 1-D kernel per input dimension
    assert per_dim.base == 3 and (per_dim.J, per_dim.K) == (3, 1) and per_dim.Z2 is None
    kern = tr.create_additive_rp_kernel(4, 3, kernel_type="Cosine", batch_kernel=False, mem_efficient=False)
    subs = kern.base_kernel.kernels
    assert all(isinstance(s.base_kernel, gk.CosineKernel) and abs(float(s.base_kernel.period_length) - 1.0) < 1e-6 for s in subs)
    assert all(abs(float(s.outputscale) - 1 / 3) < 1e-6 for s in subs)


def test_cosine_argument_reduction_in_a_float32_model():
    """kv_kernels.cuh::reduce_2pi restated in float32 (one rounding per FMA): n = round(d / 2 pi) by the 1.5 * 2^23 constant, then a
    two-term Cody-Waite subtraction of n * 2 pi.  The reduced argument stays within [-pi, pi] (to 1e-3) and its cosine is the
    cosine of d to 1.1e-7 for arguments up to 3000 -- far inside the 1e-5 of the K.V parity, with cos.approx's 2^-20.9 on top."""
    def fma(a, b, c):
        return np.float32(np.float64(a) * np.float64(b) + np.float64(c))
    rng = np.random.RandomState(0)
    d = np.concatenate([rng.rand(100000) * 60, rng.rand(100000) * 3000, np.linspace(0, 7, 1000)]).astype(np.float32)
    magic = np.float32(12582912.0)
    n = np.float32(fma(d, np.float32(0.15915494309189535), magic) - magic)
    assert np.all(n == np.rint(n)) and np.abs(n - d.astype(np.float64) / (2 * np.pi)).max() <= 0.5 + 1e-3
    r = fma(n, np.float32(-6.2831854820251465), d)
    r = fma(n, np.float32(1.7484555314695172e-07), r)
    assert np.abs(r).max() < np.pi + 1e-3        # (a hair over pi when d / 2 pi rounds the other way: harmless)
    err = np.abs(np.cos(r.astype(np.float64)) - np.cos(d.astype(np.float64)))
    assert err.max() < 1.5e-7, err.max()
    assert abs((6.2831854820251465 - 1.7484555314695172e-07) - 2 * np.pi) < 1e-14      # the split of 2 pi itself


@pytest.mark.parametrize("base", [0, 1, 2, 3])
def test_operator_diagonals_need_no_kernel_launch(base):
    """diag of K(Z, Z) is sum_j c_j for every base kernel (k(0) = 1); diag of a square K(Z1, Z2) is formed in torch"""
    g = torch.Generator().manual_seed(base)
    Z1, Z2 = torch.randn(9, 6, generator=g, dtype=torch.float64), torch.randn(9, 6, generator=g, dtype=torch.float64)
    c = torch.rand(3, generator=g, dtype=torch.float64) + 0.1
    sym = lazy.RPAdditiveLazyTensor(Z1, None, c, 3, 2, base=base)
    np.testing.assert_allclose(sym.diag().numpy(), np.full(9, float(c.sum())), rtol=1e-12)
    rect = lazy.RPAdditiveLazyTensor(Z1, Z2, c, 3, 2, base=base)
    want = np.diag(orc.additive_rbf_dense(Z1.numpy(), Z2.numpy(), c.numpy(), 3, 2, base=base))
    np.testing.assert_allclose(rect.diag().numpy(), want, rtol=1e-12)
    assert rect.shape == (9, 9) and rect._transpose_nonbatch().Z1 is Z2 and rect._rebuild(*rect.representation()).base == base

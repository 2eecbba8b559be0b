"""Projection matrices for randomly-projected additive GPs: the inputs of the K.V path.

Same API and random-number consumption as the reference's `rp.py` (so that a seeded run draws the same W):
  gen_rp         -> rp.py:10-32   (d x k matrix; the caller concatenates J of them, training_routines.py:137,144-145)
  space_equally  -> rp.py:220-268 (diversified DPA-GP directions; numpy RNG when d >= J, gradient descent otherwise)
Golden fixtures generated from the reference itself: tests/golden/gen_rp.npz, tests/golden/space_equally.npz.
The spherical t-design / Riesz-energy / ELM / PCA helpers of the reference are outside the hot path.
"""
import math

import numpy as np
import torch

RP_DISTRIBUTIONS = ("gaussian", "sphere", "very-sparse", "bernoulli", "uniform")


def gen_rp(d, k, dist="gaussian"):
    """Random projection matrix with input dimension d and output dimension k (columns are directions)."""
    root_k = math.sqrt(k)
    if dist == "gaussian":
        return torch.randn(d, k) / root_k
    if dist == "sphere":
        W = torch.randn(d, k)
        W = W / torch.norm(W, p=2, dim=0, keepdim=True)
        return W * math.sqrt(d) / root_k  # a uniform unit vector has per-coordinate variance 1/d
    if dist == "very-sparse":
        tail = 1.0 / (2.0 * math.sqrt(d))
        draws = torch.distributions.Categorical(torch.tensor([tail, 1.0 - 2.0 * tail, tail])).sample(torch.Size([d, k]))
        return (draws - 1).to(dtype=torch.float)
    if dist == "bernoulli":
        return (torch.bernoulli(torch.rand(d, k)) * 2 - 1) / root_k
    if dist == "uniform":
        return (torch.rand(d, k) * 2 - 1) / root_k * math.sqrt(3)  # U(-1,1) has variance 1/3
    raise ValueError("Not a valid RP distribution")


def _pairwise_cos4(P):
    """sum_{a != b} cos^4(angle(p_a, p_b)) -- sign invariant, differentiable."""
    unit = P / P.pow(2).sum(dim=1, keepdim=True).sqrt()
    cos = unit @ unit.t() - torch.eye(P.shape[0], dtype=P.dtype)
    return cos.pow(4).sum()


def space_equally(P, lr, niter):
    """Spread the J rows of P (J x d) apart.  Returns (new P with unit rows, final loss or None).

    d >= J: J orthonormal directions by modified Gram-Schmidt on fresh numpy Gaussian draws (the input values are
            ignored, only its shape/dtype are used -- the reference does the same, rp.py:224-239).
    d <  J: `niter` plain gradient steps of size `lr` on the pairwise cos^4 energy, then row normalisation.
    """
    J, d = P.shape
    if d >= J:
        basis = []
        for _ in range(J):
            v = np.random.randn(d)
            v = v / np.linalg.norm(v)
            for u in basis:
                v = v - v.dot(u) * u
            if basis:
                v = v / np.linalg.norm(v)
            basis.append(v)
        out = torch.from_numpy(np.vstack(basis)).to(P)
        out.requires_grad = False
        return out, None

    P = P.detach().clone().requires_grad_(True)
    for _ in range(niter):
        loss = _pairwise_cos4(P)
        (grad,) = torch.autograd.grad(loss, P)
        with torch.no_grad():
            P -= lr * grad
    final = _pairwise_cos4(P).detach().reshape(1, 1)
    with torch.no_grad():
        P /= P.pow(2).sum(dim=1, keepdim=True).sqrt()
    P.requires_grad = False
    return P, final

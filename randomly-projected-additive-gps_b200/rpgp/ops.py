"""Tensor-level operations of the K.V path on top of the C ABI (rpgp/_lib.py -> librpgp.so).

The additive-RBF normal form of SURVEY.md §0 is the only thing the kernels know about:

    K[i, i'] = sum_{j<J} c_j * exp(-1/2 * sum_{m<K} (Z1[i, jK+m] - Z2[i', jK+m])^2)

`kmatmul(Z1, Z2, c, J, K, V)` and `kdense(Z1, Z2, c, J, K)` are differentiable (torch.autograd.Function pairs, the
plug-in shape the reference uses for its own custom arithmetic -- GAMFunction, memory_efficient_gam_kernel.py:5-59);
`quad_form_grads(...)` is the direct `_quad_form_derivative(L, R)` the CG-based MLL backward calls.
float32 tensors run the fused sm_100a kernels; float64 tensors (the reference's --double path,
training_routines.py:481) run the un-tiled FP64 kernels.  CUDA only -- there is no CPU fallback.
"""
import torch

from . import _lib
from . import dist as rdist

import os

LN2 = 0.6931471805599453
# symmetric products K(Z,Z).V of at least this many rows go to the tensor-core kernel that evaluates every kernel value once
# (rpgp_mvm_sym_f32); RPGP_SYM=0 keeps everything on the SIMT forward kernel
SYM_MIN_ROWS = int(os.environ.get("RPGP_SYM_MIN_ROWS", "1024"))
SYM_ENABLED = os.environ.get("RPGP_SYM", "1") != "0"


def _use_sym(p1, p2, t, row_range):
    return (SYM_ENABLED and p2 is p1 and row_range is None and p1.n >= SYM_MIN_ROWS
            and _lib.mvm_sym_supported(p1.lay, min(t, 16)))


class Packed:
    """Packed, pre-scaled copy of natural coordinates Z (n x J*K) and of the weights c, as the fused kernels want them."""

    __slots__ = ("zp", "lay", "n")

    def __init__(self, Z, J, K, base=0):
        self.lay = _lib.plan_layout(J, K, base)
        self.zp = _lib.pack_coords(Z.detach().contiguous(), self.lay)
        self.n = Z.shape[0]


def pack_weights(c, lay):
    return _lib.pack_log2c(c.detach().reshape(-1), lay)


def unpack_coord_grad(dzp, lay):
    """(nchunks, n, CP) gradient w.r.t. packed scaled coordinates -> (n, J*K) gradient w.r.t. natural coordinates."""
    nch, n, CP = dzp.shape
    J, K, G, KP = lay.J, lay.K, lay.G, lay.KP
    g = dzp[:, :, :G * KP].reshape(nch, n, G, KP)[..., :K]          # (nch, n, G, K)
    g = g.permute(1, 0, 2, 3).reshape(n, nch * G, K)[:, :J, :]       # groups are chunk-major
    return g.reshape(n, J * K) * _lib.coord_scale()


def _expand_c(c, J, like):
    c = torch.as_tensor(c, dtype=like.dtype, device=like.device).reshape(-1)
    if c.numel() == 1 and J > 1:
        c = c.expand(J)
    if c.numel() != J:
        raise ValueError("outputscale vector has %d entries, expected J=%d" % (c.numel(), J))
    return c.contiguous()


def _check_operands(Z1, Z2, J, K):
    if Z1.dim() != 2 or Z2.dim() != 2:
        raise ValueError("coordinates must be 2-D (n x J*K); batch mode is not supported (as in GAMFunction)")
    if Z1.shape[1] != J * K or Z2.shape[1] != J * K:
        raise ValueError("Dimension mismatch")  # memory_efficient_gam_kernel.py:15-16
    _lib.require_cuda(Z1, Z2)
    if Z1.dtype != Z2.dtype or Z1.dtype not in (torch.float32, torch.float64):
        raise TypeError("coordinates must both be float32 or float64")


# ----------------------------------------------------------------------------------------------------------------------
# raw (non-differentiable) calls
# ----------------------------------------------------------------------------------------------------------------------
def kmv_raw(Z1, Z2, c, J, K, V, packed1=None, packed2=None, nlc=None, row_range=None, base=0):
    """K(Z1[rows], Z2) @ V without autograd.  `packed*` / `nlc` let the caller reuse the packed operands across the
    many products of one CG solve (Z^ is computed once per step)."""
    c = _expand_c(c, J, Z1)
    if Z1.dtype == torch.float64:
        z1 = Z1 if row_range is None else Z1[row_range[0]:row_range[1]]
        return _lib.mvm_fwd_f64(z1, Z2, c, J, K, V.to(torch.float64), base=base)
    p1 = packed1 or Packed(Z1, J, K, base)
    p2 = packed2 or (p1 if Z2 is Z1 else Packed(Z2, J, K, base))
    nlc = nlc if nlc is not None else pack_weights(c, p1.lay)
    V = V.contiguous().float()
    if _use_sym(p1, p2, V.shape[1], row_range):
        return _lib.mvm_sym(p1.zp, p1.lay, nlc, V)
    return _lib.mvm_fwd(p1.zp, p2.zp, p1.lay, nlc, V, row_range=row_range)


def quad_form_grads(Z1, Z2, c, J, K, L, R, symmetric, packed1=None, packed2=None, nlc=None, row_range=None, base=0):
    """Gradients of sum_col L[:,col]^T K(Z1,Z2) R[:,col].

    symmetric (Z2 is Z1): returns (dZ, None, dc) with dZ the TOTAL derivative (both roles), for rows `row_range`
    (all rows by default) -- the caller all-gathers row blocks and all-reduces dc across ranks.
    otherwise: returns (dZ1, dZ2, dc).
    """
    c = _expand_c(c, J, Z1)
    if Z1.dtype == torch.float64:
        L, R = L.to(torch.float64), R.to(torch.float64)
        if symmetric:
            rr = row_range or (0, Z1.shape[0])
            z1 = Z1[rr[0]:rr[1]]
            dA, g = _lib.quad_bwd_f64(z1, Z1, c, J, K, L[rr[0]:rr[1]], R, base=base)
            dB, _ = _lib.quad_bwd_f64(z1, Z1, c, J, K, R[rr[0]:rr[1]], L, base=base)
            return dA + dB, None, g / c
        dZ1, g = _lib.quad_bwd_f64(Z1, Z2, c, J, K, L, R, base=base)
        dZ2, _ = _lib.quad_bwd_f64(Z2, Z1, c, J, K, R, L, base=base)
        return dZ1, dZ2, g / c
    p1 = packed1 or Packed(Z1, J, K, base)
    lay = p1.lay
    nlc = nlc if nlc is not None else pack_weights(c, lay)
    L, R = L.contiguous().float(), R.contiguous().float()
    if symmetric:
        dzp, g = _lib.quad_bwd(p1.zp, p1.zp, lay, nlc, L, R, symmetric=True, row_range=row_range)
        return unpack_coord_grad(dzp, lay), None, g[:J] / c
    p2 = packed2 or Packed(Z2, J, K, base)
    dzp1, g = _lib.quad_bwd(p1.zp, p2.zp, lay, nlc, L, R, symmetric=False)
    dzp2, _ = _lib.quad_bwd(p2.zp, p1.zp, lay, nlc, R, L, symmetric=False)
    return unpack_coord_grad(dzp1, lay), unpack_coord_grad(dzp2, lay), g[:J] / c


def kernel_rows_raw(Zr, Z2, c, J, K, base=0):
    return _lib.kernel_rows(Zr, Z2, _expand_c(c, J, Z2), J, K, base)


# ----------------------------------------------------------------------------------------------------------------------
# differentiable wrappers
# ----------------------------------------------------------------------------------------------------------------------
class _KMatmul(torch.autograd.Function):
    """out = K(Z1, Z2) @ V; backward = quadratic-form derivative with L = grad_out, R = V, and K^T @ grad_out."""

    @staticmethod
    def forward(ctx, Z1, Z2, c, V, J, K, symmetric, base=0):
        _check_operands(Z1, Z2, J, K)
        c = _expand_c(c, J, Z1)
        ctx.J, ctx.K, ctx.symmetric, ctx.base = J, K, symmetric, base
        ctx.save_for_backward(Z1, Z2, c, V)
        Vc = V.to(Z1.dtype)
        return kmv_raw(Z1, Z1 if symmetric else Z2, c, J, K, Vc, base=base).to(V.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        Z1, Z2, c, V = ctx.saved_tensors
        J, K, sym, base = ctx.J, ctx.K, ctx.symmetric, ctx.base
        gZ1 = gZ2 = gc = gV = None
        g = grad_out.contiguous()
        need_k = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        if need_k:
            if sym:
                dZ, _, dc = quad_form_grads(Z1, Z1, c, J, K, g, V, symmetric=True, base=base)
                gZ1, gc = dZ.to(Z1.dtype), dc.to(c.dtype)
            else:
                dZ1, dZ2, dc = quad_form_grads(Z1, Z2, c, J, K, g, V, symmetric=False, base=base)
                gZ1, gZ2, gc = dZ1.to(Z1.dtype), dZ2.to(Z2.dtype), dc.to(c.dtype)
        if ctx.needs_input_grad[3]:
            gV = kmv_raw(Z1 if sym else Z2, Z1, c, J, K, g.to(Z1.dtype), base=base).to(V.dtype)  # K^T g
        return gZ1, gZ2, gc, gV, None, None, None, None


def kmatmul(Z1, Z2, c, J, K, V, base=0):
    """Differentiable K(Z1, Z2) @ V.  Pass the same tensor object for Z1 and Z2 to use the symmetric kernels."""
    squeeze = V.dim() == 1
    V2 = V.unsqueeze(-1) if squeeze else V
    sym = Z2 is Z1
    out = _KMatmul.apply(Z1, Z1 if sym else Z2, torch.as_tensor(c, device=Z1.device), V2, J, K, sym, base)
    return out.squeeze(-1) if squeeze else out


class _KDense(torch.autograd.Function):
    """Dense K(Z1, Z2) (small n: Cholesky path, tests, `evaluate()`); backward through the same row-gradient kernels
    with L = grad, R = identity, processed in column blocks."""

    @staticmethod
    def forward(ctx, Z1, Z2, c, J, K, symmetric, base=0):
        _check_operands(Z1, Z2, J, K)
        c = _expand_c(c, J, Z1)
        ctx.J, ctx.K, ctx.symmetric, ctx.base = J, K, symmetric, base
        ctx.save_for_backward(Z1, Z2, c)
        return kernel_rows_raw(Z1.detach().contiguous(), (Z1 if symmetric else Z2).detach().contiguous(), c, J, K, base)

    @staticmethod
    def backward(ctx, G):
        Z1, Z2, c = ctx.saved_tensors
        J, K, sym, base = ctx.J, ctx.K, ctx.symmetric, ctx.base
        m, n = G.shape
        Zb = Z1 if sym else Z2
        dZ1 = torch.zeros_like(Z1)
        dZ2 = torch.zeros_like(Zb)
        dc = torch.zeros_like(c)
        blk = 16
        eye = torch.eye(n, dtype=Z1.dtype, device=Z1.device)
        for c0 in range(0, n, blk):
            c1 = min(n, c0 + blk)
            L = G[:, c0:c1].contiguous()
            R = eye[:, c0:c1].contiguous()
            a, b, g = quad_form_grads(Z1, Zb, c, J, K, L, R, symmetric=False, base=base)
            dZ1 += a
            dZ2 += b
            dc += g
        if sym:
            return dZ1 + dZ2, None, dc, None, None, None, None
        return dZ1, dZ2, dc, None, None, None, None


def kdense(Z1, Z2, c, J, K, base=0):
    sym = Z2 is Z1
    return _KDense.apply(Z1, Z1 if sym else Z2, torch.as_tensor(c, device=Z1.device), J, K, sym, base)


# ----------------------------------------------------------------------------------------------------------------------
# the projection Z = ((x * pre_inv) W^T) * post_inv as one library call with an explicit vector-Jacobian product
# ----------------------------------------------------------------------------------------------------------------------
class _Project(torch.autograd.Function):
    """Z[i, q] = post_inv[q] * sum_k x[i, k] * pre_inv[k] * W[q, k]  (scaled_projection_kernel.py:21-27: prescale = pre_inv = 1 / l over
    the d inputs, postscale = post_inv = 1 / l over the J*K outputs).  Forward: rpgp_project2_f32 (tcgen05, 3xTF32); backward:
    rpgp_project_bwd_f32 gives dW' = dZ^T x, from which the gradients of W, pre_inv and post_inv are element-wise products of
    (J*K x d) matrices; the gradient with respect to x (never needed by the training path) is a dense product."""

    @staticmethod
    def forward(ctx, x, W, pre_inv, post_inv):
        JK, d = W.shape
        lay = _lib.plan_layout(JK, 1)          # only the natural output is produced here: any layout of the right width will do
        _, Z = _lib.project2(x.detach(), W.detach(), None if pre_inv is None else pre_inv.detach(),
                             None if post_inv is None else post_inv.detach(), lay, packed=False, natural=True)
        ctx.save_for_backward(x, W, pre_inv, post_inv)
        return Z

    @staticmethod
    def backward(ctx, gZ):
        x, W, pre_inv, post_inv = ctx.saved_tensors
        gx = gW = gpre = gpost = None
        need_w = ctx.needs_input_grad[1] or (pre_inv is not None and ctx.needs_input_grad[2]) or \
            (post_inv is not None and ctx.needs_input_grad[3])
        pre = None if pre_inv is None else pre_inv.reshape(1, -1)
        post = None if post_inv is None else post_inv.reshape(-1, 1)
        if need_w:
            dWp = _lib.project_bwd(x.detach(), gZ.contiguous().float())           # (JK x d) = dZ^T x
            Wd = W.detach()
            if ctx.needs_input_grad[1]:
                gW = dWp if pre is None else dWp * pre
                gW = gW if post is None else gW * post
            if pre_inv is not None and ctx.needs_input_grad[2]:
                t = dWp * Wd
                gpre = (t if post is None else t * post).sum(0).reshape(pre_inv.shape)
            if post_inv is not None and ctx.needs_input_grad[3]:
                t = dWp * Wd
                gpost = (t if pre is None else t * pre).sum(1).reshape(post_inv.shape)
        if ctx.needs_input_grad[0]:
            Weff = W.detach()
            Weff = Weff if pre is None else Weff * pre
            Weff = Weff if post is None else Weff * post
            gx = gZ @ Weff
        return gx, gW, gpre, gpost


def project(x, W, pre_inv=None, post_inv=None):
    """Differentiable projection through the C ABI (float32 CUDA, 2-D); see _Project."""
    return _Project.apply(x, W, pre_inv, post_inv)


def can_project(x, W):
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(-1) == 1 and W.dtype == torch.float32
            and W.dim() == 2 and x.shape[0] > 0)


# ----------------------------------------------------------------------------------------------------------------------
# row-partitioned products (multi-GPU): each rank computes its row block, NCCL all-gather rebuilds the product
# ----------------------------------------------------------------------------------------------------------------------
def kmv_partitioned(Z, c, J, K, V, packed=None, nlc=None, base=0):
    """Symmetric K(Z,Z) @ V with the rows of K split over the ranks of the default process group."""
    part = rdist.partition(Z.shape[0])
    if part.world == 1:
        return kmv_raw(Z, Z, c, J, K, V, packed1=packed, packed2=packed, nlc=nlc, base=base)
    if Z.dtype == torch.float32:
        p = packed or Packed(Z, J, K, base)
        if _use_sym(p, p, V.shape[1], None):
            # symmetric tensor-core kernel: rank r owns the unique block pairs of its share of the 128-row blocks and
            # produces partial sums for ALL rows -> the exchange is an all-reduce (sum) instead of an all-gather
            nblocks = (p.n + 127) // 128
            per = (nblocks + part.world - 1) // part.world
            b0, b1 = min(nblocks, part.rank * per), min(nblocks, (part.rank + 1) * per)
            w = nlc if nlc is not None else pack_weights(_expand_c(c, J, Z), p.lay)
            out = _lib.mvm_sym(p.zp, p.lay, w, V.contiguous().float(), block_range=(b0, b1))
            return rdist.all_reduce_sum(out)
    blk = kmv_raw(Z, Z, c, J, K, V, packed1=packed, packed2=packed, nlc=nlc, row_range=(part.r0, part.r1), base=base)
    return rdist.all_gather_rows(blk, part)


RECT_MIN_ROWS_PER_RANK = 64


def kmv_rect_partitioned(Z1, Z2, c, J, K, V, packed1=None, packed2=None, nlc=None, base=0):
    """Rectangular K(Z1, Z2) @ V (prediction: test rows x training columns) with the ROWS of Z1 split over the ranks and one
    all-gather of the row blocks ("test rows partitioned the same way; all-gather of n* means").  Every rank must call
    it with the same replicated operands; small products (fewer than RECT_MIN_ROWS_PER_RANK rows per rank) stay replicated."""
    part = rdist.partition(Z1.shape[0])
    if part.world == 1 or Z1.shape[0] < RECT_MIN_ROWS_PER_RANK * part.world:
        return kmv_raw(Z1, Z2, c, J, K, V, packed1=packed1, packed2=packed2, nlc=nlc, base=base)
    blk = kmv_raw(Z1, Z2, c, J, K, V, packed1=packed1, packed2=packed2, nlc=nlc, row_range=(part.r0, part.r1), base=base)
    return rdist.all_gather_rows(blk.contiguous(), part)

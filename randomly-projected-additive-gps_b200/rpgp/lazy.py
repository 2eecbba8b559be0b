"""Lazy operators: what the reference's kernels hand to the solver instead of a dense matrix.

`RPAdditiveLazyTensor` is the drop-in for the object graph the reference builds out of GPyTorch pieces
(SumLazyTensor of J KeOpsLazyTensor / LazyEvaluatedKernelTensor; plug-in pattern at gp_models/kernels/imq_kernel.py:51-58):
it exposes the methods the CG / Lanczos MLL and the prediction strategy call on any LazyTensor -- `_matmul`, `_size`,
`_transpose_nonbatch`, `diag`, row access, `_quad_form_derivative(left, right)`, `representation`, `evaluate` -- and
routes all arithmetic to the fused sm_100a kernels (rpgp/ops.py).  `AddedDiagLazyTensor` is K + sigma_n^2 I (what the
likelihood adds), `DenseLazyTensor` wraps an explicit matrix (predictive covariances).
"""
import torch

from . import dist as rdist
from . import ops
from .solver.linear_cg import linear_cg
from .solver.preconditioner import PivCholPreconditioner, pivoted_cholesky


def _settings():
    from .gp import settings
    return settings


class LazyTensor:
    """Minimal LazyTensor protocol (the subset of gpytorch.lazy.LazyTensor the reference's call sites reach)."""

    def _size(self):
        raise NotImplementedError

    @property
    def shape(self):
        return self._size()

    def size(self, dim=None):
        s = self._size()
        return s if dim is None else s[dim]

    def dim(self):
        return len(self._size())

    # ---- algebra -----------------------------------------------------------------------------------------------------
    def _matmul(self, rhs):
        """raw product, no autograd graph (solver inner loop)"""
        raise NotImplementedError

    def matmul(self, rhs):
        raise NotImplementedError

    def __matmul__(self, rhs):
        return self.matmul(rhs)

    def add_diag(self, diag):
        return AddedDiagLazyTensor(self, diag)

    def add_jitter(self, jitter_val=1e-3):
        return self.add_diag(torch.tensor(jitter_val, dtype=self.dtype, device=self.device))

    def evaluate_kernel(self):
        return self

    def numpy(self):
        return self.evaluate().detach().cpu().numpy()

    def __getitem__(self, index):
        if isinstance(index, tuple) and len(index) == 2 and isinstance(index[1], slice) and index[1] == slice(None):
            index = index[0]
        if isinstance(index, tuple):
            return self.evaluate()[index]
        if isinstance(index, int):
            return self.rows(torch.tensor([index], device=self.device))[0]
        if isinstance(index, slice):
            index = torch.arange(*index.indices(self.shape[0]), device=self.device)
        return self.rows(torch.as_tensor(index, device=self.device))


def base_function(base, sq):
    """k(d^2) of the base kernels (include/rpgp.h rpgp_base_kernel), torch, elementwise"""
    if base == 1:
        q = (3.0 * sq).clamp_min(0).sqrt()
        return (1.0 + q) * torch.exp(-q)
    if base == 2:
        return (sq + 1.0).rsqrt()
    if base == 3:
        return torch.cos(sq.clamp_min(0).sqrt())
    return torch.exp(-0.5 * sq)


class RPAdditiveLazyTensor(LazyTensor):
    """K(Z1, Z2) = sum_j c_j exp(-1/2 |Z1[:, group j] - Z2[:, group j]|^2), never materialised.

    Z1 (m x J*K), Z2 (n x J*K) or None for the symmetric K(Z1, Z1), c (J,) positive weights.
    All learnable hyper-parameters of every reference family enter only through Z and c, so
    `representation()` is (Z1[, Z2], c) and autograd chains the rest.
    """

    def __init__(self, Z1, Z2, c, J, K, base=0):
        if Z1.dim() != 2:
            raise ValueError("RPAdditiveLazyTensor does not support batch mode (neither does GAMFunction)")
        if Z2 is not None and Z2.shape[-1] != Z1.shape[-1]:
            raise ValueError("Dimension mismatch")
        if Z1.shape[-1] != J * K:
            raise ValueError("Dimension mismatch")
        self.Z1 = Z1
        self.Z2 = None if (Z2 is None or Z2 is Z1) else Z2
        self.c = c if torch.is_tensor(c) else torch.as_tensor(c, dtype=Z1.dtype, device=Z1.device)
        if self.c.dim() == 0 or self.c.numel() == 1:
            self.c = self.c.reshape(1).expand(J)
        self.J, self.K = int(J), int(K)
        self.base = int(base)       # base kernel of every group: 0 RBF, 1 Matern-1.5, 2 inverse multiquadric (include/rpgp.h)
        self._packed1 = self._packed2 = self._nlc = None

    # ---- structure ---------------------------------------------------------------------------------------------------
    @property
    def symmetric(self):
        return self.Z2 is None

    @property
    def dtype(self):
        return self.Z1.dtype

    @property
    def device(self):
        return self.Z1.device

    def _size(self):
        return torch.Size((self.Z1.shape[0], (self.Z1 if self.symmetric else self.Z2).shape[0]))

    def representation(self):
        return (self.Z1, self.c) if self.symmetric else (self.Z1, self.Z2, self.c)

    def _rebuild(self, *rep):
        if self.symmetric:
            return RPAdditiveLazyTensor(rep[0], None, rep[1], self.J, self.K, self.base)
        return RPAdditiveLazyTensor(rep[0], rep[1], rep[2], self.J, self.K, self.base)

    def detach(self):
        return self._rebuild(*[r.detach() for r in self.representation()])

    def _transpose_nonbatch(self):
        if self.symmetric:
            return self
        return RPAdditiveLazyTensor(self.Z2, self.Z1, self.c, self.J, self.K, self.base)

    def transpose(self, a=-2, b=-1):
        return self._transpose_nonbatch()

    t = _transpose_nonbatch

    def scale(self, s):
        """outputscale * K  (ScaleKernel)"""
        return RPAdditiveLazyTensor(self.Z1, self.Z2, self.c * s, self.J, self.K, self.base)

    __mul__ = scale
    __rmul__ = scale

    @staticmethod
    def sum(ops_list):
        """Sum of additive-RBF operators over the same points = one operator with concatenated groups
        (AdditiveKernel / SumLazyTensor); narrower groups are zero-padded to the widest K."""
        ops_list = list(ops_list)
        first = ops_list[0]
        if any(o.base != first.base for o in ops_list):
            raise NotImplementedError("a sum of different base kernels does not lower to one fused operator")
        Kmax = max(o.K for o in ops_list)
        sym = all(o.symmetric for o in ops_list)

        def widen(Z, o):
            if o.K == Kmax:
                return Z
            Zg = Z.reshape(Z.shape[0], o.J, o.K)
            pad = Zg.new_zeros(Z.shape[0], o.J, Kmax - o.K)
            return torch.cat([Zg, pad], dim=-1).reshape(Z.shape[0], o.J * Kmax)

        Z1 = torch.cat([widen(o.Z1, o) for o in ops_list], dim=-1)
        Z2 = None if sym else torch.cat([widen(o.Z1 if o.symmetric else o.Z2, o) for o in ops_list], dim=-1)
        c = torch.cat([o.c.reshape(-1) for o in ops_list])
        return RPAdditiveLazyTensor(Z1, Z2, c, sum(o.J for o in ops_list), Kmax, first.base)

    def __add__(self, other):
        if isinstance(other, RPAdditiveLazyTensor):
            return RPAdditiveLazyTensor.sum([self, other])
        return NotImplemented

    @staticmethod
    def product(ops_list):
        """Product of single-group RBF operators on disjoint coordinates = one RBF on the concatenated coordinates
        (ProductKernel of 1-D RBFs, polynomial_projection_kernels.py:88-92)."""
        ops_list = list(ops_list)
        if any(o.J != 1 or o.base != 0 for o in ops_list):
            raise NotImplementedError("only products of single-group RBF kernels lower to the fused operator")
        sym = all(o.symmetric for o in ops_list)
        Z1 = torch.cat([o.Z1 for o in ops_list], dim=-1)
        Z2 = None if sym else torch.cat([o.Z1 if o.symmetric else o.Z2 for o in ops_list], dim=-1)
        c = ops_list[0].c
        for o in ops_list[1:]:
            c = c * o.c
        return RPAdditiveLazyTensor(Z1, Z2, c, 1, Z1.shape[-1])

    # ---- packed operands are built once and reused by every product of a solve ----------------------------------------
    def _prepare(self):
        if self.dtype != torch.float32:
            return
        if self._packed1 is None:
            self._packed1 = ops.Packed(self.Z1, self.J, self.K, self.base)
            self._packed2 = self._packed1 if self.symmetric else ops.Packed(self.Z2, self.J, self.K, self.base)
            self._nlc = ops.pack_weights(self.c, self._packed1.lay)

    # ---- arithmetic ------------------------------------------------------------------------------------------------------
    def _matmul(self, rhs):
        squeeze = rhs.dim() == 1
        V = rhs.unsqueeze(-1) if squeeze else rhs
        self._prepare()
        with torch.no_grad():
            if self.symmetric and rdist.world_size() > 1:
                out = ops.kmv_partitioned(self.Z1.detach(), self.c.detach(), self.J, self.K, V.detach(),
                                          packed=self._packed1, nlc=self._nlc, base=self.base)
            elif not self.symmetric and rdist.world_size() > 1:
                out = ops.kmv_rect_partitioned(self.Z1.detach(), self.Z2.detach(), self.c.detach(), self.J, self.K, V.detach(),
                                               packed1=self._packed1, packed2=self._packed2, nlc=self._nlc, base=self.base)
            else:
                Z2 = self.Z1 if self.symmetric else self.Z2
                out = ops.kmv_raw(self.Z1.detach(), Z2.detach(), self.c.detach(), self.J, self.K, V.detach(),
                                  packed1=self._packed1, packed2=self._packed2, nlc=self._nlc, base=self.base)
        out = out.to(rhs.dtype)
        return out.squeeze(-1) if squeeze else out

    def matmul(self, rhs):
        """differentiable product"""
        Z2 = self.Z1 if self.symmetric else self.Z2
        return ops.kmatmul(self.Z1, Z2, self.c, self.J, self.K, rhs, base=self.base)

    def _quad_form_derivative(self, left_vecs, right_vecs):
        """d/d(representation) of sum_col left[:,col]^T K right[:,col]  (no graph; used by the MLL backward)."""
        self._prepare()
        L = left_vecs.unsqueeze(-1) if left_vecs.dim() == 1 else left_vecs
        R = right_vecs.unsqueeze(-1) if right_vecs.dim() == 1 else right_vecs
        with torch.no_grad():
            if self.symmetric:
                part = rdist.partition(self.Z1.shape[0])
                rr = None if part.world == 1 else (part.r0, part.r1)
                dZ, _, dc = ops.quad_form_grads(self.Z1.detach(), self.Z1.detach(), self.c.detach(), self.J, self.K,
                                                L.detach(), R.detach(), True, packed1=self._packed1, nlc=self._nlc,
                                                row_range=rr, base=self.base)
                if part.world > 1:
                    dZ = rdist.all_gather_rows(dZ.contiguous(), part)
                    dc = rdist.all_reduce_sum(dc.contiguous())
                return dZ.to(self.Z1.dtype), dc.to(self.c.dtype)
            dZ1, dZ2, dc = ops.quad_form_grads(self.Z1.detach(), self.Z2.detach(), self.c.detach(), self.J, self.K,
                                               L.detach(), R.detach(), False, packed1=self._packed1,
                                               packed2=self._packed2, nlc=self._nlc, base=self.base)
            return dZ1.to(self.Z1.dtype), dZ2.to(self.Z2.dtype), dc.to(self.c.dtype)

    def diag(self):
        if self.symmetric:
            return self.c.sum().expand(self.Z1.shape[0])
        if self.Z1.shape != self.Z2.shape:
            raise RuntimeError("diag of a non-square operator")
        d = (self.Z1 - self.Z2).reshape(self.Z1.shape[0], self.J, self.K)
        return (self.c * base_function(self.base, (d * d).sum(-1))).sum(-1)

    _approx_diag = diag

    def rows(self, index):
        """K[index, :] as a dense (len(index) x n) tensor (pivoted-Cholesky row fetch); no graph."""
        Z2 = self.Z1 if self.symmetric else self.Z2
        with torch.no_grad():
            return ops.kernel_rows_raw(self.Z1.detach()[index].contiguous(), Z2.detach().contiguous(), self.c.detach(),
                                       self.J, self.K, self.base)

    def evaluate(self):
        """dense matrix (small n only), differentiable"""
        Z2 = self.Z1 if self.symmetric else self.Z2
        return ops.kdense(self.Z1, Z2, self.c, self.J, self.K, self.base)


class DenseLazyTensor(LazyTensor):
    """An explicit matrix behind the LazyTensor protocol (gpytorch NonLazyTensor)."""

    def __init__(self, tensor):
        self.tensor = tensor

    dtype = property(lambda self: self.tensor.dtype)
    device = property(lambda self: self.tensor.device)

    def _size(self):
        return self.tensor.shape

    def representation(self):
        return (self.tensor,)

    def _rebuild(self, *rep):
        return DenseLazyTensor(rep[0])

    def _transpose_nonbatch(self):
        return DenseLazyTensor(self.tensor.transpose(-1, -2))

    def _matmul(self, rhs):
        return self.tensor.detach() @ rhs

    def matmul(self, rhs):
        return self.tensor @ rhs

    def _quad_form_derivative(self, left_vecs, right_vecs):
        return (left_vecs @ right_vecs.transpose(-1, -2),)

    def diag(self):
        return self.tensor.diagonal(dim1=-2, dim2=-1)

    _approx_diag = diag

    def rows(self, index):
        return self.tensor.detach()[index]

    def evaluate(self):
        return self.tensor


class ZeroLazyTensor(LazyTensor):
    """what prediction returns for the covariance under settings.skip_posterior_variances (training_routines.py:551)"""

    def __init__(self, *sizes, dtype=None, device=None):
        self._sizes, self._dtype, self._device = torch.Size(sizes), dtype, device

    dtype = property(lambda self: self._dtype)
    device = property(lambda self: self._device)

    def _size(self):
        return self._sizes

    def _matmul(self, rhs):
        return torch.zeros(self._sizes[:-1] + rhs.shape[-1:], dtype=rhs.dtype, device=rhs.device)

    matmul = _matmul

    def diag(self):
        return torch.zeros(self._sizes[0], dtype=self._dtype, device=self._device)

    def evaluate(self):
        return torch.zeros(self._sizes, dtype=self._dtype, device=self._device)

    def add_diag(self, diag):
        return DenseLazyTensor(torch.diag_embed(diag.reshape(-1).expand(self._sizes[0]).to(self._dtype)))


class PredictiveCovarLazyTensor(LazyTensor):
    """K** - K*X A KX*, the exact-GP predictive covariance kept lazy (gpytorch DefaultPredictionStrategy.exact_predictive_covar;
    reached from training_routines.py:551-575).  A is K^-1 applied by multi-right-hand-side CG (exact), or, under
    settings.fast_pred_var, the cached Lanczos root W W^T ~= K^-1 (LOVE; gp_experiment_runner.py:235,327).  Nothing of
    size n x n* is ever formed: `diag()` works through the test points in batches, `_matmul` costs one solve (exact) or two
    thin products (LOVE)."""

    def __init__(self, test_test, cross, train_covar, root=None):
        self.test_test, self.cross, self.train_covar = test_test, cross, train_covar
        self.root = root                      # n x r, K^-1 ~= root root^T
        self._cross_root = None               # n* x r = K*X root

    dtype = property(lambda self: self.cross.dtype)
    device = property(lambda self: self.cross.device)

    def _size(self):
        m = self.cross.shape[0]
        return torch.Size((m, m))

    def _transpose_nonbatch(self):
        return self

    def _cross_times_root(self):
        if self._cross_root is None:
            self._cross_root = self.cross._matmul(self.root)
        return self._cross_root

    def _matmul(self, rhs):
        squeeze = rhs.dim() == 1
        V = rhs.unsqueeze(-1) if squeeze else rhs
        out = self.test_test._matmul(V)
        if self.root is not None:
            cr = self._cross_times_root()
            out = out - cr @ (cr.t() @ V)
        else:
            out = out - self.cross._matmul(self.train_covar.inv_matmul(self.cross._transpose_nonbatch()._matmul(V)))
        return out.squeeze(-1) if squeeze else out

    matmul = _matmul

    def rows(self, index):
        index = torch.as_tensor(index, device=self.device).reshape(-1)
        out = self.test_test.rows(index)
        if self.root is not None:
            cr = self._cross_times_root()
            return out - cr[index] @ cr.t()
        c_rows = self.cross.rows(index)                                    # b x n
        solves = self.train_covar.inv_matmul(c_rows.t().contiguous())       # n x b
        return out - self.cross._matmul(solves).t()

    def diag(self):
        prior = self.test_test.diag()
        if self.root is not None:
            cr = self._cross_times_root()
            return prior - (cr * cr).sum(-1)
        m = self.cross.shape[0]
        out = torch.empty(m, dtype=self.dtype, device=self.device)
        step = max(1, int(_settings().variance_batch_size.value()))
        for i0 in range(0, m, step):
            idx = torch.arange(i0, min(m, i0 + step), device=self.device)
            c_rows = self.cross.rows(idx)                                   # b x n   (K(X*_b, X))
            solves = self.train_covar.inv_matmul(c_rows.t().contiguous())   # n x b   (one multi-RHS CG solve)
            out[idx] = prior[idx] - (c_rows.t() * solves).sum(0)
        return out

    _approx_diag = diag

    def evaluate(self):
        m = self.cross.shape[0]
        return self._matmul(torch.eye(m, dtype=self.dtype, device=self.device))


class AddedDiagLazyTensor(LazyTensor):
    """K + sigma^2 I: the operator CG actually solves with (GaussianLikelihood adds the noise, never the kernel)."""

    def __init__(self, base, noise):
        self.base = base
        self.noise = noise if torch.is_tensor(noise) else torch.tensor(float(noise), dtype=base.dtype, device=base.device)
        self._precond = None
        self._precond_built = False

    dtype = property(lambda self: self.base.dtype)
    device = property(lambda self: self.base.device)

    def _size(self):
        return self.base._size()

    def _noise_scalar(self):
        return self.noise.reshape(-1)[0]

    def _matmul(self, rhs):
        return self.base._matmul(rhs) + self._noise_scalar().detach().to(rhs.dtype) * rhs

    def matmul(self, rhs):
        return self.base.matmul(rhs) + self._noise_scalar() * rhs

    def representation(self):
        if isinstance(self.base, PredictiveCovarLazyTensor):
            return (self.noise,)
        return tuple(self.base.representation()) + (self.noise,)

    def _rebuild(self, *rep):
        if isinstance(self.base, PredictiveCovarLazyTensor):
            return AddedDiagLazyTensor(self.base, rep[-1])
        return AddedDiagLazyTensor(self.base._rebuild(*rep[:-1]), rep[-1])

    def _quad_form_derivative(self, left_vecs, right_vecs):
        if isinstance(self.base, PredictiveCovarLazyTensor):      # evaluation only: no hyper-parameter gradients through it
            return ((left_vecs * right_vecs).sum().reshape(self.noise.shape).to(self.noise.dtype),)
        base = self.base._quad_form_derivative(left_vecs, right_vecs)
        dnoise = (left_vecs * right_vecs).sum().reshape(self.noise.shape).to(self.noise.dtype)
        return tuple(base) + (dnoise,)

    def diag(self):
        return self.base.diag() + self._noise_scalar()

    def rows(self, index):
        r = self.base.rows(index).clone()
        idx = torch.as_tensor(index, device=r.device).reshape(-1)
        r[torch.arange(idx.numel(), device=r.device), idx] += self._noise_scalar().detach()
        return r

    def evaluate(self):
        K = self.base.evaluate()
        return K + self._noise_scalar() * torch.eye(K.shape[-1], dtype=K.dtype, device=K.device)

    def add_diag(self, diag):
        return AddedDiagLazyTensor(self.base, self.noise + diag)

    # ---- preconditioner (n >= min_preconditioning_size only) -------------------------------------------------------------
    def _preconditioner(self):
        s = _settings()
        if self._precond_built:
            return self._precond
        self._precond_built = True
        n = self.shape[-1]
        rank = s.max_preconditioner_size.value()
        if rank == 0 or n < s.min_preconditioning_size.value():
            self._precond = None
        else:
            with torch.no_grad():
                L = pivoted_cholesky(self.base, rank)
                self._precond = PivCholPreconditioner(L, self._noise_scalar().detach().to(L.dtype))
        return self._precond

    # ---- solves ------------------------------------------------------------------------------------------------------------
    def _use_cholesky(self):
        s = _settings()
        return (not s.fast_computations.solves.on()) or self.shape[-1] <= s.max_cholesky_size.value()

    def _solve(self, rhs, preconditioner, num_tridiag=0, tolerance=None):
        s = _settings()
        tol = s.cg_tolerance.value() if tolerance is None else tolerance
        return linear_cg(self._matmul, rhs, n_tridiag=num_tridiag, max_iter=s.max_cg_iterations.value(),
                         max_tridiag_iter=s.max_lanczos_quadrature_iterations.value(), tolerance=tol,
                         preconditioner=None if preconditioner is None else preconditioner.solve)

    def inv_matmul(self, rhs):
        """K^-1 rhs without autograd (prediction caches; uses eval_cg_tolerance like gpytorch in eval mode)."""
        s = _settings()
        with torch.no_grad():
            if self._use_cholesky():
                from .solver.inv_quad_logdet import psd_safe_cholesky
                Lc = psd_safe_cholesky(self.evaluate().detach())
                return torch.cholesky_solve(rhs if rhs.dim() > 1 else rhs.unsqueeze(-1), Lc).reshape(rhs.shape)
            return self._solve(rhs, self._preconditioner(), tolerance=s.eval_cg_tolerance.value())

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False):
        """(rhs^T K^-1 rhs summed over columns, log|K|), differentiable w.r.t. representation() and rhs."""
        from .solver.inv_quad_logdet import inv_quad_logdet
        return inv_quad_logdet(self, inv_quad_rhs, logdet)

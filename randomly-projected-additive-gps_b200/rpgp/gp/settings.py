"""Solver settings as context managers, with GPyTorch's names and defaults -- the knobs the reference's CLI sets at
gp_experiment_runner.py:324-332 and synthetic_test_script.py:122-123.

    with settings.cg_tolerance(0.002), settings.eval_cg_tolerance(0.001), settings.max_cg_iterations(10_000): ...
"""


class _Value:
    _default = None
    _value = None

    def __init__(self, value):
        self._new = value

    @classmethod
    def value(cls):
        return cls._default if cls._value is None else cls._value

    @classmethod
    def _set(cls, v):
        cls._value = v

    def __enter__(self):
        self._prev = self.__class__._value
        self.__class__._set(self._new)
        return self

    def __exit__(self, *exc):
        self.__class__._set(self._prev)
        return False


class _Flag:
    _default = False
    _state = None

    def __init__(self, state=True):
        self._new = bool(state)

    @classmethod
    def on(cls):
        return cls._default if cls._state is None else cls._state

    @classmethod
    def off(cls):
        return not cls.on()

    def __enter__(self):
        self._prev = self.__class__._state
        self.__class__._state = self._new
        return self

    def __exit__(self, *exc):
        self.__class__._state = self._prev
        return False


class cg_tolerance(_Value):
    """mean residual norm at which training-mode CG stops (gpytorch default 1; reference CLI default 0.05)"""
    _default = 1.0


class eval_cg_tolerance(_Value):
    _default = 0.01


class max_cg_iterations(_Value):
    _default = 1000


class cg_convergence_lag(_Value):
    """How many iterations late the host may learn that CG has converged (not a GPyTorch setting).  GPyTorch reads the residual norm
    on the host after EVERY iteration (a device synchronisation per iteration).  -1 (default): 0 for products of >= 2^22 kernel-matrix
    rows x right-hand sides (the iteration is device-bound: an extra iteration costs more than the synchronisation), 2 below that
    (launch-bound sizes such as BASELINE configs[0]: the residual norm is copied to pinned memory asynchronously and read two
    iterations later, so the host keeps launching; the solve runs up to two iterations past GPyTorch's stopping point -- never
    fewer -- and reports the residual it actually reached).  0 restores GPyTorch's iteration counts exactly."""
    _default = -1


class max_cholesky_size(_Value):
    _default = 800


class max_preconditioner_size(_Value):
    _default = 15


class min_preconditioning_size(_Value):
    _default = 2000


class num_trace_samples(_Value):
    _default = 10


class max_lanczos_quadrature_iterations(_Value):
    _default = 20


class checkpoint_kernel(_Value):
    """accepted for CLI compatibility (gp_experiment_runner.py:330); the fused kernels never materialise K, so the
    row-chunk size has no effect"""
    _default = 0


class max_root_decomposition_size(_Value):
    """rank of the Lanczos root of K^-1 that fast_pred_var (LOVE) caches; gpytorch default 100"""
    _default = 100


class max_dense_predictive_size(_Value):
    """n* x n entries up to which the predictive covariance is formed densely (exact, what gpytorch's default strategy
    does); above it -- or under fast_pred_var -- the covariance stays lazy (rpgp/lazy.py PredictiveCovarLazyTensor)"""
    _default = 1 << 25


class variance_batch_size(_Value):
    """test points per multi-right-hand-side solve when exact predictive variances are computed lazily"""
    _default = 64


class skip_posterior_variances(_Flag):
    _default = False


class skip_logdet_forward(_Flag):
    _default = False


class memory_efficient(_Flag):
    _default = False


class fast_pred_var(_Flag):
    _default = False


class use_toeplitz(_Flag):
    _default = True


class debug(_Flag):
    _default = True


class _FastSolves(_Flag):
    _default = True


class _FastLogDet(_Flag):
    _default = True


class _FastCovarRoot(_Flag):
    _default = True


class fast_computations:
    """fast_computations(covar_root_decomposition, log_prob, solves); --use_chol turns all three off
    (gp_experiment_runner.py:325)"""
    covar_root_decomposition = _FastCovarRoot
    log_prob = _FastLogDet
    solves = _FastSolves

    def __init__(self, covar_root_decomposition=True, log_prob=True, solves=True):
        self._ctx = [_FastCovarRoot(covar_root_decomposition), _FastLogDet(log_prob), _FastSolves(solves)]

    def __enter__(self):
        for c in self._ctx:
            c.__enter__()
        return self

    def __exit__(self, *exc):
        for c in reversed(self._ctx):
            c.__exit__(*exc)
        return False


class deterministic_probes:
    """Fix the SLQ probe vectors (n x num_trace_samples) -- for reproducible MLL values in tests."""
    probe_vectors = None

    def __init__(self, probe_vectors):
        self._new = probe_vectors

    def __enter__(self):
        self._prev = deterministic_probes.probe_vectors
        deterministic_probes.probe_vectors = self._new
        return self

    def __exit__(self, *exc):
        deterministic_probes.probe_vectors = self._prev
        return False

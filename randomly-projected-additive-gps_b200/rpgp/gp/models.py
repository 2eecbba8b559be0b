"""ExactGP: prior in train mode, exact posterior in eval mode (gpytorch.models.ExactGP; gp_models/models.py:10-20).

Prediction restates DefaultPredictionStrategy: mean cache = K^-1 (y - mu) by CG with
eval_cg_tolerance (one square multi-iteration solve, t = 1), predictive mean = K(X*, X) mean_cache + mu (one rectangular
K.V), covariance = K** - K*X K^-1 KX* (n* right-hand sides) unless settings.skip_posterior_variances.  Small problems form
the covariance densely; beyond settings.max_dense_predictive_size, or under settings.fast_pred_var (LOVE: cached Lanczos root of
K^-1, gp_experiment_runner.py:235,327), it stays lazy (rpgp/lazy.py PredictiveCovarLazyTensor) and variances come out in batches.
"""
import torch

from .. import lazy
from . import settings
from .distributions import MultivariateNormal
from .module import Module


class ExactGP(Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if train_inputs is not None and torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = None if train_inputs is None else tuple(
            t.unsqueeze(-1) if t.dim() == 1 else t for t in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood
        self._mean_cache = None
        self._love_root = None
        self._cache_key = None

    def _apply(self, fn):
        if self.train_inputs is not None:
            self.train_inputs = tuple(fn(t) for t in self.train_inputs)
            self.train_targets = fn(self.train_targets)
        return super()._apply(fn)

    def train(self, mode=True):
        if mode:
            self._mean_cache = None
            self._love_root = None
        return super().train(mode)

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            if torch.is_tensor(inputs):
                inputs = (inputs,)
            self.train_inputs = tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in inputs)
        if targets is not None:
            self.train_targets = targets
        self._mean_cache = None
        self._love_root = None

    def __call__(self, *args, **kwargs):
        inputs = [a.unsqueeze(-1) if a.dim() == 1 else a for a in args]
        if self.training:
            if settings.debug.on() and self.train_inputs is not None:
                if not all(torch.equal(a, b) for a, b in zip(self.train_inputs, inputs)):
                    raise RuntimeError("You must train on the training inputs!")
            return self.forward(*inputs, **kwargs)
        if self.train_inputs is None or self.train_targets is None:
            return self.forward(*inputs, **kwargs)
        return self._predict(inputs[0])

    # ---- exact prediction ------------------------------------------------------------------------------------------------
    def _prediction_cache_key(self):
        """the prediction caches (K^-1 (y - mu), the train operator, the LOVE root) are valid for one state of the parameters:
        every in-place change -- optimizer steps, load_state_dict, `initialize`, property setters -- bumps a tensor's `_version`
        (edits through `.data` do not and are not detected)"""
        return tuple((id(p), p._version) for p in self.parameters()) + (id(self.train_inputs[0]), id(self.train_targets))

    def _predict(self, x_test):
        x_train = self.train_inputs[0]
        key = self._prediction_cache_key()
        if key != self._cache_key:              # parameters changed while the model stayed in eval mode
            self._mean_cache = None
            self._love_root = None
            self._cache_key = key
        with torch.no_grad():
            prior = self.forward(x_train)
            if self._mean_cache is None:
                train_covar = self.likelihood(prior).lazy_covariance_matrix           # K + sigma^2 I
                resid = (self.train_targets - prior.mean).unsqueeze(-1)
                self._mean_cache = train_covar.inv_matmul(resid)
                self._train_covar = train_covar
        test_mean = self.mean_module(x_test)
        cross = self.covar_module(x_test, x_train).evaluate_kernel()               # K(X*, X)
        pred_mean = test_mean + cross._matmul(self._mean_cache).squeeze(-1)
        n_test = x_test.shape[-2]
        if settings.skip_posterior_variances.on():
            covar = lazy.ZeroLazyTensor(n_test, n_test, dtype=pred_mean.dtype, device=pred_mean.device)
            return MultivariateNormal(pred_mean, covar)
        lazy_cov = settings.fast_pred_var.on() or n_test * x_train.shape[-2] > settings.max_dense_predictive_size.value()
        if lazy_cov and isinstance(cross, lazy.RPAdditiveLazyTensor):
            with torch.no_grad():
                root = None
                if settings.fast_pred_var.on():     # LOVE: Lanczos root of K^-1 started from the mean cross-covariance column
                    if self._love_root is None:
                        from ..solver.lanczos import lanczos_root_inv
                        ones = torch.full((n_test, 1), 1.0 / n_test, dtype=pred_mean.dtype, device=pred_mean.device)
                        init = cross._transpose_nonbatch()._matmul(ones)
                        rank = min(settings.max_root_decomposition_size.value(), x_train.shape[-2])
                        self._love_root = lanczos_root_inv(self._train_covar._matmul, init, rank)
                    root = self._love_root
                test_test = self.covar_module(x_test).evaluate_kernel()
                covar = lazy.PredictiveCovarLazyTensor(test_test.detach(), cross.detach(), self._train_covar, root=root)
            return MultivariateNormal(pred_mean, covar)
        with torch.no_grad():
            cross_dense = cross.evaluate().detach()                                    # n* x n
            solves = self._train_covar.inv_matmul(cross_dense.t().contiguous())       # n x n*
            test_test = self.covar_module(x_test).evaluate_kernel().evaluate().detach()
            pred_covar = test_test - cross_dense @ solves
            pred_covar = 0.5 * (pred_covar + pred_covar.t())
        return MultivariateNormal(pred_mean, lazy.DenseLazyTensor(pred_covar))

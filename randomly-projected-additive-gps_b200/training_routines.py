"""Kernel factories and the exact-GP train / evaluate routine for the randomly-projected additive families.

API mirror of the reference's training_routines.py for the kinds that sit on the K.V hot path:
  create_rp_poly_kernel      (:108-128)  `rp_poly`      -- additive_rp_J20_K1.json and siblings
  create_additive_rp_kernel  (:131-189)  `additive_rp`  -- additive_rp_{pre,post}scale_*.json, additive_spread_prescale_*.json
  create_strictly_additive_kernel / create_additive_kernel (GAM-style baselines over raw features, RBF only)
  create_exact_gp            (:325-410)  outer ScaleKernel, GaussianLikelihood with the SmoothedBoxPrior on the noise
  train_exact_gp             (:469-585)  random restarts, training, metric dictionary
Same function names, keyword arguments, defaults and `model_specs/*.json` schema ({kind, model_kwargs, train_kwargs});
the other `kind`s of the reference (full, sgpr, multi_full, duvenaud_additive, deep_rp_poly, SKI variants, ppr, cgp,
model averaging) are different model families outside the hot path and raise NotImplementedError.
"""
import copy
import json
import os
import warnings

import numpy as np
import torch

import rp
from config import model_base_path
from fitting.optimizing import mean_squared_error, train_to_convergence
from gp_models.kernels import (CustomAdditiveKernel, GeneralizedProjectionKernel, MemoryEfficientGamKernel,
                               PolynomialProjectionKernel, ScaledProjectionKernel, StrictlyAdditiveKernel)
from gp_models.models import ExactGPModel
from rpgp import gp as gpytorch
from rpgp.gp.kernels import CosineKernel, InverseMQKernel, MaternKernel, RBFKernel, ScaleKernel

SPEC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_specs")
HOT_PATH_KINDS = ("rp_poly", "additive_rp", "strictly_additive", "additive", "general_rp_poly")
REFERENCE_KINDS = ("full", "rp", "strictly_additive", "additive", "rp_poly", "deep_rp_poly", "general_rp_poly",
                   "multi_full", "duvenaud_additive", "additive_rp", "sgpr")


def load_model_spec(name_or_path):
    """Read a {kind, model_kwargs, train_kwargs} spec (gp_experiment_runner.py:258-259); bare names resolve in
    model_specs/."""
    path = name_or_path
    if not os.path.exists(path):
        path = os.path.join(SPEC_DIR, name_or_path if name_or_path.endswith(".json") else name_or_path + ".json")
    with open(path) as fh:
        return json.load(fh)


def _map_to_optim(optimizer):
    """optimizer name -> torch optimizer class"""
    table = {"adam": torch.optim.Adam, "sgd": torch.optim.SGD, "lbfgs": torch.optim.LBFGS}
    if optimizer not in table:
        raise ValueError("Unknown optimizer")
    return table[optimizer]


def _save_state_dict(model):
    """Save the state dict under a content-derived file name; returns the file name."""
    state = model.state_dict()
    fname = "model_state_dict_{}.pkl".format(hash(str(state)))
    folder = os.path.join(model_base_path, "models")
    os.makedirs(folder, exist_ok=True)
    torch.save(state, os.path.join(folder, fname))
    return fname


def _sample_from_range(num_samples, range_):
    return torch.rand(num_samples) * (range_[1] - range_[0]) + range_[0]


def _map_to_kernel(return_object, kernel_type, keops, **key_words):
    """Base-kernel lookup (reference :52-83).  `keops` is accepted and ignored: there is a single backend -- the hand-written
    kernels replace both the dense and the KeOps route of the reference.  RBF, Matern (nu = 1.5), the inverse multiquadric and the
    cosine kernel (dense-only in the reference, :76-81) are base kernels 0..3 of the fused kernels."""
    if return_object:
        cls, kwargs = _map_to_kernel(False, kernel_type, keops)
        return cls(**key_words, **kwargs)
    if kernel_type == "RBF":
        return RBFKernel, dict(**key_words)
    if kernel_type == "Matern":
        return MaternKernel, dict(nu=1.5, **key_words)
    if kernel_type == "InverseMQ":
        return InverseMQKernel, dict(**key_words)
    if kernel_type == "Cosine":
        return CosineKernel, dict(**key_words)
    raise ValueError("Unknown kernel type")


def _no_ski(ski):
    if ski:
        raise NotImplementedError("SKI specs use an interpolation approximation, not the exact K.V path")


def create_rp_poly_kernel(d, k, J, activation=None, learn_proj=False, weighted=False, kernel_type="RBF", space_proj=False,
                          init_mixin_range=(1.0, 1.0), init_lengthscale_range=(1.0, 1.0), ski=False, ski_options=None,
                          X=None, proj_dist="gaussian", keops=False):
    _no_ski(ski)
    projs = [rp.gen_rp(d, k, dist=proj_dist) for _ in range(J)]
    bs = [torch.zeros(k) for _ in range(J)]
    if space_proj:
        newW, _ = rp.space_equally(torch.cat(projs, dim=1).t(), lr=0.1, niter=5000)
        newW.requires_grad = False
        projs = [newW[i:i + 1, :].t() for i in range(J)]
    kernel_cls, kwargs = _map_to_kernel(False, kernel_type, keops)
    kernel = PolynomialProjectionKernel(J, k, d, kernel_cls, projs, bs, activation=activation, learn_proj=learn_proj,
                                        weighted=weighted, ski=ski, ski_options=ski_options, X=X, **kwargs)
    kernel.initialize(init_mixin_range, init_lengthscale_range)
    return kernel


def create_additive_rp_kernel(d, J, learn_proj=False, kernel_type="RBF", space_proj=False, prescale=False, ard=True,
                              init_lengthscale_range=(1., 1.), ski=False, ski_options=None, proj_dist="gaussian",
                              batch_kernel=True, mem_efficient=False, k=1, keops=False):
    _no_ski(ski)
    if k > 1 and (mem_efficient or batch_kernel or space_proj):
        raise ValueError("Can't have k > 1 with memory efficient GAM kernel or a batch kernel or spaced projections.")
    projs = [rp.gen_rp(d, k, dist=proj_dist) for _ in range(J)]
    if space_proj:
        newW, _ = rp.space_equally(torch.cat(projs, dim=1).t(), lr=0.1, niter=5000)
        newW.requires_grad = False
        projs = [newW[i:i + k, :].t() for i in range(0, J * k, k)]
    proj_module = torch.nn.Linear(d, J * k, bias=False)
    proj_module.weight.data = torch.cat(projs, dim=1).t().contiguous()

    def make_kernel(active_dim=None):
        kernel = _map_to_kernel(True, kernel_type, keops, active_dims=active_dim)
        if hasattr(kernel, "period_length"):      # the cosine kernel has no lengthscale (reference :150-153)
            kernel.initialize(period_length=torch.tensor([1.]))
        else:
            kernel.initialize(lengthscale=torch.tensor([1.]))
        kernel = ScaleKernel(kernel)
        kernel.initialize(outputscale=torch.tensor([1 / J]))
        return kernel

    if mem_efficient:
        if batch_kernel:
            raise ValueError("Impossible to have batch kernel and memory efficient GAM")
        if kernel_type != "RBF":
            raise ValueError("Memory efficient GAM with alternative sub-kernels not implemented yet.")
        add_kernel = MemoryEfficientGamKernel()  # lengthscale stays at softplus(0) = ln 2 and there is no 1/J (as in the reference)
    elif batch_kernel:
        add_kernel = gpytorch.kernels.AdditiveStructureKernel(make_kernel(None), J)
    else:
        add_kernel = gpytorch.kernels.AdditiveKernel(*[make_kernel(list(range(i, i + k))) for i in range(0, J * k, k)])
    if ard:
        ard_num_dims = d if prescale else J * k
        initial_ls = _sample_from_range(ard_num_dims, init_lengthscale_range)
    else:
        ard_num_dims = None
        initial_ls = _sample_from_range(1, init_lengthscale_range)
    proj_kernel = ScaledProjectionKernel(proj_module, add_kernel, prescale=prescale, ard_num_dims=ard_num_dims,
                                         learn_proj=learn_proj)
    proj_kernel.initialize(lengthscale=initial_ls)
    return proj_kernel


def create_general_rp_poly_kernel(d, degrees, learn_proj=False, weighted=False, kernel_type="RBF",
                                  init_lengthscale_range=(1.0, 1.0), init_mixin_range=(1.0, 1.0), ski=False,
                                  ski_options=None, X=None, keops=False):
    _no_ski(ski)
    out_dim = sum(degrees)
    W = torch.cat([rp.gen_rp(d, 1) for _ in range(out_dim)], dim=1).t()
    projection_module = torch.nn.Linear(d, out_dim, bias=False)
    projection_module.weight = torch.nn.Parameter(W)
    projection_module.bias = torch.nn.Parameter(torch.zeros(out_dim))
    kernel_cls, kwargs = _map_to_kernel(False, kernel_type, keops)
    kernel = GeneralizedProjectionKernel(degrees, d, kernel_cls, projection_module, learn_proj, weighted, ski, ski_options,
                                         X=X, **kwargs)
    kernel.initialize(init_mixin_range, init_lengthscale_range)
    return kernel


def create_strictly_additive_kernel(d, weighted=False, kernel_type="RBF", init_lengthscale_range=(1.0, 1.0),
                                    init_mixin_range=(1.0, 1.0), ski=False, ski_options=None, X=None, keops=False,
                                    memory_efficient=False):
    _no_ski(ski)
    if kernel_type == "RBF" and memory_efficient:
        kernel = MemoryEfficientGamKernel(ard_num_dims=d)
        kernel.initialize(lengthscale=_sample_from_range(d, init_lengthscale_range))
        return kernel
    kernel_cls, kwargs = _map_to_kernel(False, kernel_type, keops)
    kernel = StrictlyAdditiveKernel(d, kernel_cls, weighted, ski=ski, ski_options=ski_options, X=X, **kwargs)
    kernel.initialize(init_mixin_range, init_lengthscale_range)
    return kernel


def create_additive_kernel(d, groups, weighted=False, kernel_type="RBF", init_lengthscale_range=(1.0, 1.0),
                           init_mixin_range=(1.0, 1.0), ski=False, ski_options=None, X=None, keops=False):
    _no_ski(ski)
    kernel_cls, kwargs = _map_to_kernel(False, kernel_type, keops)
    kernel = CustomAdditiveKernel(groups, d, kernel_cls, weighted, ski=ski, ski_options=ski_options, X=X, **kwargs)
    kernel.initialize(init_mixin_range, init_lengthscale_range)
    return kernel


def create_exact_gp(trainX, trainY, kind, devices=("cpu",), **kwargs):
    """Exact GP with the kernel structure `kind`, an outer ScaleKernel, and a Gaussian likelihood."""
    [n, d] = trainX.shape
    if kind not in REFERENCE_KINDS:
        raise ValueError("Unknown kernel structure type {}".format(kind))
    if kind not in HOT_PATH_KINDS:
        raise NotImplementedError("kind '%s' is a different model family, outside the K.V hot path (SURVEY.md §2)" % kind)

    noise_prior_ = gpytorch.priors.SmoothedBoxPrior(1e-4, 10, sigma=0.01) if kwargs.pop("noise_prior") else None
    likelihood = gpytorch.likelihoods.GaussianLikelihood(noise_prior=noise_prior_)
    likelihood.noise = _sample_from_range(1, kwargs.pop("init_noise_range", [1.0, 1.0]))
    kwargs.pop("grid_size", None)
    kwargs.pop("grid_ratio", None)
    if kind == "rp_poly":
        kernel = create_rp_poly_kernel(d, X=trainX, **kwargs)
    elif kind == "general_rp_poly":
        kernel = create_general_rp_poly_kernel(d, X=trainX, **kwargs)
    elif kind == "additive_rp":
        kernel = create_additive_rp_kernel(d, **kwargs)
    elif kind == "strictly_additive":
        kernel = create_strictly_additive_kernel(d, X=trainX, **kwargs)
    else:  # additive
        kernel = create_additive_kernel(d, X=trainX, **kwargs)

    kernel = gpytorch.kernels.ScaleKernel(kernel)
    if len(devices) > 1:
        # the reference wraps the kernel in MultiDeviceKernel (:407-408); here the rows of K are partitioned over the
        # ranks of torch.distributed (one process per GPU, torchrun) inside the operator itself -- nothing to wrap.
        warnings.warn("devices=%s ignored: multi-GPU runs use one process per GPU (torchrun); see rpgp/dist.py" % (devices,))
    model = ExactGPModel(trainX, trainY, likelihood, kernel)
    return model, likelihood


def train_exact_gp(trainX, trainY, testX, testY, kind, model_kwargs, train_kwargs, devices=("cpu",),
                   skip_posterior_variances=False, skip_random_restart=False, evaluate_on_train=True,
                   output_device=None, record_pred_unc=False, double=False):
    """Create and train an exact GP with the given options; returns (metrics, test predictive mean on CPU, model)."""
    model_kwargs = copy.copy(model_kwargs)
    train_kwargs = copy.copy(train_kwargs)
    d = trainX.shape[-1]
    devices = [torch.device(device) for device in devices]
    output_device = devices[0] if output_device is None else torch.device(output_device)
    type_ = torch.double if double else torch.float
    trainX, trainY = trainX.to(output_device, type_), trainY.to(output_device, type_)
    testX, testY = testX.to(output_device, type_), testY.to(output_device, type_)

    for key, v in list(model_kwargs.items()):  # "d" stands for the data dimension (e.g. J = "d")
        if isinstance(v, str) and v == "d":
            model_kwargs[key] = d

    random_restarts = train_kwargs.pop("random_restarts", 1)
    init_iters = train_kwargs.pop("init_iters", 20)
    optimizer_ = _map_to_optim(train_kwargs.pop("optimizer"))
    rr_check_conv = train_kwargs.pop("rr_check_conv", False)
    initial_train_kwargs = copy.copy(train_kwargs)
    initial_train_kwargs["max_iter"] = init_iters
    initial_train_kwargs["check_conv"] = rr_check_conv

    def fresh():
        model, likelihood = create_exact_gp(trainX, trainY, kind, devices=devices, **model_kwargs)
        model = model.to(output_device, type_)
        return model, likelihood, gpytorch.mlls.ExactMarginalLogLikelihood(likelihood, model)

    if not skip_random_restart:
        best, best_loss = None, np.inf
        for _ in range(random_restarts):  # truncated training from several initialisations, keep the best
            model, likelihood, mll = fresh()
            train_to_convergence(model, trainX, trainY, optimizer=optimizer_, objective=mll, isloss=False,
                                 **initial_train_kwargs)
            model.train()
            loss = -mll(model(trainX), trainY).item()
            if loss < best_loss:
                best_loss, best = loss, (model, likelihood, mll)
        model, likelihood, mll = best
    else:
        model, likelihood, mll = fresh()

    # default warning filter, as the reference (training_routines.py:535): repeated warnings from one location are recorded once
    with warnings.catch_warnings(record=True) as w:
        trained_epochs = train_to_convergence(model, trainX, trainY, optimizer=optimizer_, objective=mll, isloss=False,
                                              **train_kwargs)

    model.eval()
    likelihood.eval()
    mll.eval()
    model_metrics = dict()
    model_metrics["trained_epochs"] = trained_epochs
    with torch.no_grad():
        model.train()  # the prior, for the MLL of the training data
        likelihood.train()
        model_metrics["prior_train_nmll"] = -mll(model(trainX), trainY).item()
        with gpytorch.settings.skip_posterior_variances(skip_posterior_variances):
            model.eval()  # now posterior distributions
            likelihood.eval()
            if evaluate_on_train:
                train_outputs = model(trainX)
                model_metrics["train_mse"] = mean_squared_error(train_outputs.mean, trainY)
            with warnings.catch_warnings(record=True) as w2:
                test_outputs = model(testX)
                pred_mean = test_outputs.mean
            if not skip_posterior_variances:
                if evaluate_on_train:
                    model_metrics["train_nll"] = -mll(train_outputs, trainY).item()
                model_metrics["test_nll"] = -mll(test_outputs, testY).item()
                distro = likelihood(test_outputs)
                lower, upper = distro.confidence_region()
                frac = ((testY > lower) * (testY < upper)).to(torch.float).mean().item()
                model_metrics["test_pred_frac_in_cr"] = frac
                if record_pred_unc:
                    model_metrics["test_pred_z_score"] = (testY - distro.mean) / distro.stddev
    model_metrics["training_warnings"] = len(w)
    model_metrics["testing_warning"] = "" if len(w2) == 0 else w2[-1].message
    model_metrics["state_dict_file"] = _save_state_dict(model)
    return model_metrics, pred_mean.to("cpu", torch.float), model

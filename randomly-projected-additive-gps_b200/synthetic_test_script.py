"""Synthetic-function benchmark of the reference (synthetic_test_script.py): learning curves of exact GPs on d-dimensional test
functions, trained with CG through the fused K.V path -- BASELINE configs[0] is one point of such a curve (n = 2 000, d = 10,
additive target, additive_rp_J20_K1).

Mirrors the reference's target functions (:20-75), `benchmark_on_n_pts` (:78-147: data law `rand * 4 - 2`, targets + 0.01 noise,
normalisation by the HOLD-OUT statistics, solver settings cg_tolerance 1e-3 / eval 5e-4 / 10 000 CG iterations, Adam through
`train_to_convergence`), `benchmark_algo_on_func` (:150-168) and the model constructors that lower to the fused operator (:183-202).
Differences: nothing runs at import time (the reference executes its configuration block on import, :211-232 -- here it is `main()`);
the device is an argument instead of the hard-coded 'cuda:7' (:18); the Matern-2.5 baselines (`create_bl_model`, `create_rp_model`,
:169-181) use nu = 1.5, the Matern order the fused kernels provide, and say so.
"""
import argparse
import gc
import json
from math import pi

import numpy as np
import torch

from fitting.optimizing import mean_squared_error, train_to_convergence
from gp_models import ExactGPModel, RPPolyKernel
from rpgp.gp import settings as gp_set
from rpgp.gp.kernels import MaternKernel, RBFKernel, ScaleKernel
from rpgp.gp.likelihoods import GaussianLikelihood
from rpgp.gp.mlls import ExactMarginalLogLikelihood
from training_routines import create_additive_rp_kernel, create_strictly_additive_kernel

device = "cuda:0"


# ---- target functions (x: n x d) ------------------------------------------------------------------------------------------------
def unimodal_d_dim(x):
    return torch.exp(-torch.norm(x, dim=1) ** 2)


def bimodal_d_dim(x):
    one = torch.ones(1, x.shape[1]).to(x)
    return torch.exp(-torch.norm(x + one, dim=1)) + torch.exp(-torch.norm(x - one, dim=1))


def multimodal_d_dim(x):
    d = x.shape[1]
    centers = 2.0 * torch.eye(d).to(x) - 1.0          # centre i: +1 in coordinate i, -1 elsewhere
    return sum(torch.exp(-torch.norm(x - centers[i:i + 1], dim=1)) for i in range(d))


def leading_dim(x):
    return bimodal_d_dim(x[:, 1:]) * 0.4 + torch.sin(x[:, 0] * pi)


def one_dim(x):
    return torch.sin(x[:, 0] * pi)


def half_relevant(x):
    return unimodal_d_dim(x[:, :x.shape[1] // 2])


def nonseparable(x):
    return x.prod(dim=-1)


def additive(x):
    return torch.sin(x).sum(dim=-1)


def non_additive(x):
    """continuous XOR: mixture of Gaussians at +-1.4 e_i"""
    d = x.shape[1]
    centers = torch.cat([torch.eye(d), -torch.eye(d)]).to(x) * 1.4          # (2d, d)
    sq = (x.unsqueeze(1) - centers.unsqueeze(0)).pow(2).sum(dim=-1)        # (n, 2d)
    return torch.exp(-3 * sq).sum(dim=1)


TARGETS = {f.__name__: f for f in (unimodal_d_dim, bimodal_d_dim, multimodal_d_dim, leading_dim, one_dim, half_relevant,
                                   nonseparable, additive, non_additive)}


# ---- benchmark loops --------------------------------------------------------------------------------------------------------------
def benchmark_on_n_pts(n_pts, create_model_func, target_func, ho_x, ho_y, fit=True, repeats=3, max_iter=1000, return_model=False,
                       verbose=0, checkpoint=True, print_freq=1, use_chol=False, device=None, **kwargs):
    """`repeats` independent training sets of n_pts points; returns (hold-out MSEs, models, mlls).  Inputs and targets of both sets
    are normalised by the hold-out statistics (the reference's choice, :99-107)."""
    dev = torch.device(device or globals()["device"])
    dims = ho_x.shape[1]
    mx, sx, my, sy = ho_x.mean(dim=0), ho_x.std(dim=0), ho_y.mean(), ho_y.std()
    test_x, test_y = ((ho_x - mx) / sx).to(dev), ((ho_y - my) / sy).to(dev)
    rep_mses, models, mlls = [], [], []
    fast = not use_chol
    for i in range(repeats):
        data = torch.rand(n_pts, dims) * 4 - 2
        y = target_func(data) + torch.randn(n_pts) * 0.01
        data, y = ((data - mx) / sx).to(dev), ((y - my) / sy).to(dev)
        model = create_model_func(data, y, **kwargs).to(dev)
        mll = ExactMarginalLogLikelihood(model.likelihood, model)
        with gp_set.fast_computations(fast, fast, fast), gp_set.max_cg_iterations(10_000), gp_set.cg_tolerance(0.001), \
                gp_set.eval_cg_tolerance(0.0005), gp_set.memory_efficient(True):
            if fit:
                train_to_convergence(model, data, y, torch.optim.Adam, objective=mll, checkpoint=checkpoint, max_iter=max_iter,
                                     print_freq=print_freq, verbose=verbose)
            model.eval()
            with torch.no_grad():
                mse = mean_squared_error(model(test_x).mean, test_y)
        print(i, mse)
        rep_mses.append(mse)
        if return_model:
            models.append(model)
            mlls.append(mll)
    torch.cuda.empty_cache()
    gc.collect()
    return rep_mses, models, mlls


SIZES = (10, 20, 40, 80, 160, 320, 640, 1280, 2560, 5120, 10240)


def benchmark_algo_on_func(create_model_func, target_func, dims=6, max_pts=2560, fit=True, repeats=3, start_after=0, use_chol=False,
                           progress_file=None, **kwargs):
    """learning curve: mean hold-out RMSE over `repeats` for the training-set sizes of SIZES in (start_after, max_pts]; the 4 000
    hold-out points are drawn once (:153-154)"""
    rmses = []
    ho_x = torch.rand(4000, dims) * 4 - 2
    ho_y = target_func(ho_x)
    for n_pts in SIZES:
        if n_pts <= start_after:
            continue
        if n_pts > max_pts:
            break
        print("n_pts={}".format(n_pts))
        rep_mses, _, _ = benchmark_on_n_pts(n_pts, create_model_func, target_func, ho_x, ho_y, fit=fit, repeats=repeats,
                                            use_chol=use_chol, **kwargs)
        rmses.append(float(np.mean(np.sqrt(rep_mses))))
        if progress_file:
            json.dump(rmses, open(progress_file, "w"))
    return rmses


# ---- models -------------------------------------------------------------------------------------------------------------------------
def create_bl_model(data, y):
    """baseline: one Matern kernel over all dimensions (nu = 1.5 here, 2.5 in the reference :169-172; d <= 32)"""
    return ExactGPModel(data, y, GaussianLikelihood(), ScaleKernel(MaternKernel(nu=1.5)))


def create_rp_model(data, y, proj_ratio=1):
    """diversified 1-D projections with Matern components (nu = 1.5 here, 2.5 in the reference :175-180)"""
    d = data.shape[1]
    kernel = ScaleKernel(RPPolyKernel(round(proj_ratio * d), 1, d, MaternKernel, nu=1.5, weighted=True, space_proj=True))
    return ExactGPModel(data, y, GaussianLikelihood(), kernel)


def create_poly_rp_model(data, y, J, k):
    d = data.shape[1]
    kernel = ScaleKernel(RPPolyKernel(J, k, d, RBFKernel, weighted=True, space_proj=True))
    return ExactGPModel(data, y, GaussianLikelihood(), kernel)


def create_dpa_gp_ard_model(data, y, J):
    d = data.shape[1]
    kernel = ScaleKernel(create_additive_rp_kernel(d, J, learn_proj=False, kernel_type="RBF", space_proj=True, prescale=True,
                                                   batch_kernel=False, ard=True, proj_dist="sphere", mem_efficient=True))
    return ExactGPModel(data, y, GaussianLikelihood(), kernel)


def create_gam_model(data, y):
    d = data.shape[1]
    kernel = ScaleKernel(create_strictly_additive_kernel(d, False, "RBF", memory_efficient=True))
    return ExactGPModel(data, y, GaussianLikelihood(), kernel)


MODELS = {"bl": create_bl_model, "rp": create_rp_model, "poly_rp": create_poly_rp_model, "dpa_ard": create_dpa_gp_ard_model,
          "gam": create_gam_model}


def main(argv=None):
    """the reference's configuration block (:211-232) as a command line: one learning curve per model, written as JSON"""
    global device
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--func", default="additive", choices=sorted(TARGETS))
    ap.add_argument("--models", nargs="+", default=["gam", "dpa_ard"], choices=sorted(MODELS))
    ap.add_argument("--dims", type=int, default=6)
    ap.add_argument("--min_pts", type=int, default=600)
    ap.add_argument("--max_pts", type=int, default=12000)
    ap.add_argument("--repeats", type=int, default=15)
    ap.add_argument("--max_iter", type=int, default=1000)
    ap.add_argument("--use_chol", action="store_true")
    ap.add_argument("--J", type=int, default=None, help="projections for dpa_ard / poly_rp (default: dims)")
    ap.add_argument("--k", type=int, default=1, help="coordinates per projection for poly_rp")
    ap.add_argument("--device", default=device)
    ap.add_argument("-o", "--output", default="synthetic_experiment.json")
    args = ap.parse_args(argv)
    device = args.device
    out = {}
    for name in args.models:
        extra = {}
        if name == "dpa_ard":
            extra = {"J": args.J or args.dims}
        elif name == "poly_rp":
            extra = {"J": args.J or args.dims, "k": args.k}
        out[name] = benchmark_algo_on_func(MODELS[name], TARGETS[args.func], dims=args.dims, start_after=args.min_pts,
                                           max_pts=args.max_pts, repeats=args.repeats, use_chol=args.use_chol, max_iter=args.max_iter,
                                           **extra)
        json.dump(out, open(args.output, "w"))
    return out


if __name__ == "__main__":
    main()

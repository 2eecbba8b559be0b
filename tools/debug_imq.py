"""Debugging aid: the inverse-MQ training test step by step (where do non-finite values first appear?)."""
import os, sys, warnings
import numpy as np, torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "randomly-projected-additive-gps_b200")); sys.path.insert(0, ROOT)
import training_routines as tr
from rpgp.gp import settings
DEV = torch.device("cuda:0")
X = torch.rand(900, 5, generator=torch.Generator().manual_seed(1)) * 4 - 2
y = torch.sin(X).sum(-1); y = (y - y.mean()) / y.std()
Xt = torch.rand(200, 5, generator=torch.Generator().manual_seed(2)) * 4 - 2
yt = torch.sin(Xt).sum(-1); yt = (yt - torch.sin(X).sum(-1).mean()) / torch.sin(X).sum(-1).std()
spec = tr.load_model_spec("additive_rp_J20_K1")
spec["model_kwargs"]["kernel_type"] = "InverseMQ"
spec["train_kwargs"].update(max_iter=int(os.environ.get("ITERS", "30")), check_conv=False, verbose=1)
torch.manual_seed(3); np.random.seed(3)
lagv = int(os.environ.get("LAG", "-1"))
with settings.cg_tolerance(0.01), settings.eval_cg_tolerance(1e-3), settings.cg_convergence_lag(lagv), warnings.catch_warnings():
    warnings.simplefilter("ignore")
    spec["train_kwargs"]["verbose"] = 0
    metrics, pred, model = tr.train_exact_gp(X, y, Xt, yt, spec["kind"], spec["model_kwargs"], spec["train_kwargs"],
                                             devices=("cuda:0",), skip_random_restart=True, skip_posterior_variances=True)
    print("rmse", float(((pred - yt) ** 2).mean().sqrt()), "noise", float(model.likelihood.noise))
    model.eval(); model.likelihood.eval()
    import importlib
    lmod = importlib.import_module('rpgp.solver.linear_cg')
    for rep in range(3):
        with torch.no_grad():
            lmod.STATS["iterations"] = 0; lmod.STATS["solves"] = 0
            out = model(Xt.to(DEV))
            cov = out.lazy_covariance_matrix.evaluate() if hasattr(out, "lazy_covariance_matrix") else out.covariance_matrix
            ev = torch.linalg.eigvalsh(cov.double())
            print("rep", rep, "cov finite", bool(torch.isfinite(cov).all()), "eig min/max", float(ev.min()), float(ev.max()),
                  "diag min", float(cov.diagonal().min()), "cg iterations", lmod.STATS["iterations"], "solves", lmod.STATS["solves"])
        model._mean_cache = None

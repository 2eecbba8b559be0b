"""Generate golden fixtures from the UNMODIFIED reference, in the build container (where /root/reference exists).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

What is executed from the reference:
  * rp.py                 imported as a module (pure torch/numpy/scipy): gen_rp (rp.py:10-32), space_equally (rp.py:220-268)
  * postprocess_inverse_mq  the function gp_models/kernels/imq_kernel.py:8-9 (same slicing + exec; `python make_golden.py imq`
                          writes only imq.npz): squared distance -> inverse-multiquadric kernel value
  * GAMFunction           the class body of gp_models/kernels/memory_efficient_gam_kernel.py:5-59, extracted by source
                          slicing + exec because the module header does `import gpytorch`, which is not installed here
                          (SURVEY.md §8c).  No reference source is copied into the repo: only its OUTPUTS are stored.
The GPU box has no /root/reference; tests only read the .npz files written here.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_rp():
    spec = importlib.util.spec_from_file_location("ref_rp", os.path.join(REF, "rp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_gam_function():
    src = open(os.path.join(REF, "gp_models/kernels/memory_efficient_gam_kernel.py")).read()
    start = src.index("class GAMFunction")
    end = src.index("class MemoryEfficientGamKernel")
    ns = {"torch": torch}
    exec(compile(src[start:end], "ref_GAMFunction", "exec"), ns)
    return ns["GAMFunction"]


def load_postprocess_inverse_mq():
    src = open(os.path.join(REF, "gp_models/kernels/imq_kernel.py")).read()
    start = src.index("def postprocess_inverse_mq")
    end = src.index("class InverseMQKernel")
    ns = {"torch": torch}
    exec(compile(src[start:end], "ref_postprocess_inverse_mq", "exec"), ns)
    return ns["postprocess_inverse_mq"]


def main_imq():
    """InverseMQKernel.forward (imq_kernel.py:17-22): inputs divided by the lengthscale, squared distance, the reference's own
    post-processing function; the squared distance (gpytorch covar_dist in the reference) is formed with torch.cdist here."""
    post = load_postprocess_inverse_mq()
    out = {}
    sq = torch.linspace(0, 50, 101, dtype=torch.double)
    out["grid_sq"] = sq.numpy().copy()
    out["grid_k"] = post(sq.clone()).numpy()
    for idx, (n, m, d) in enumerate([(9, 7, 1), (16, 12, 3), (5, 11, 6)]):
        g = torch.Generator().manual_seed(300 + idx)
        x1 = torch.randn(n, d, generator=g, dtype=torch.double) * 2
        x2 = torch.randn(m, d, generator=g, dtype=torch.double) * 2
        ls = torch.rand(1, d, generator=g, dtype=torch.double) + 0.5
        dist = torch.cdist(x1 / ls, x2 / ls) ** 2
        out["c%d_x1" % idx], out["c%d_x2" % idx], out["c%d_ls" % idx] = x1.numpy(), x2.numpy(), ls.numpy()
        out["c%d_K" % idx] = post(dist.clone()).numpy()
    np.savez(os.path.join(OUT, "imq.npz"), **out)
    print("wrote imq.npz")


def main_synth():
    """the target functions of synthetic_test_script.py:20-75 (source slicing + exec: the module imports gpytorch / matplotlib and
    runs an experiment at import time); `python make_golden.py synth` writes only synthetic_targets.npz"""
    src = open(os.path.join(REF, "synthetic_test_script.py")).read()
    start = src.index("def unimodal_d_dim")
    end = src.index("def benchmark_on_n_pts")
    from math import pi, sqrt
    ns = {"torch": torch, "pi": pi, "sqrt": sqrt, "np": np}
    exec(compile(src[start:end], "ref_synthetic_targets", "exec"), ns)
    g = torch.Generator().manual_seed(77)
    x = torch.rand(9, 6, generator=g) * 4 - 2
    out = {"x": x.numpy()}
    for name in ["unimodal_d_dim", "bimodal_d_dim", "multimodal_d_dim", "leading_dim", "one_dim", "half_relevant", "nonseparable",
                 "additive", "non_additive"]:
        out[name] = ns[name](x.clone()).numpy()
    np.savez(os.path.join(OUT, "synthetic_targets.npz"), **out)
    print("wrote synthetic_targets.npz")


def main():
    rp = load_rp()
    GAM = load_gam_function()

    # ---- G4/G5: the reference's own known-answer inputs (test.py:640-680), outputs from the reference's GAMFunction
    x1 = torch.tensor([[0., 2., 4.], [3., 4.3, 2.], [6.2, 1.2, 2.2]], dtype=torch.double, requires_grad=True)
    x2 = torch.tensor([[3., 2., 1.], [5.3, 2.1, 7.1]], dtype=torch.double, requires_grad=True)
    raw = torch.tensor([1., 3., 2.], dtype=torch.double, requires_grad=True)
    ls = torch.nn.functional.softplus(raw)
    Kg = GAM.apply(x1, x2, ls)
    Kg.sum().backward()
    np.savez(os.path.join(OUT, "gam_g4_g5.npz"), x1=x1.detach().numpy(), x2=x2.detach().numpy(),
             raw_lengthscale=raw.detach().numpy(), lengthscale=ls.detach().numpy(), K=Kg.detach().numpy(),
             dx1=x1.grad.numpy(), dx2=x2.grad.numpy(), draw=raw.grad.numpy())

    # ---- seeded random GAMFunction cases, FP64 and FP32, with a random upstream gradient (the quadratic form S = L R^T)
    cases = {}
    for idx, (n, m, d, ard) in enumerate([(17, 23, 5, True), (64, 40, 20, True), (33, 33, 26, False), (128, 96, 20, True)]):
        g = torch.Generator().manual_seed(100 + idx)
        a = torch.randn(n, d, generator=g, dtype=torch.double) * 1.5
        b = torch.randn(m, d, generator=g, dtype=torch.double) * 1.5
        ell = torch.rand(d if ard else 1, generator=g, dtype=torch.double) + 0.5
        L = torch.randn(n, 3, generator=g, dtype=torch.double)
        R = torch.randn(m, 3, generator=g, dtype=torch.double)
        for dt, tag in ((torch.double, "f64"), (torch.float, "f32")):
            a_, b_, e_ = (t.to(dt).clone().requires_grad_(True) for t in (a, b, ell))
            Kc = GAM.apply(a_, b_, e_)
            (Kc * (L.to(dt) @ R.to(dt).t())).sum().backward()
            pre = "c%d_%s_" % (idx, tag)
            cases.update({pre + "x1": a_.detach().numpy(), pre + "x2": b_.detach().numpy(),
                          pre + "ell": e_.detach().numpy(), pre + "L": L.to(dt).numpy(), pre + "R": R.to(dt).numpy(),
                          pre + "K": Kc.detach().numpy(), pre + "dx1": a_.grad.numpy(), pre + "dx2": b_.grad.numpy(),
                          pre + "dell": e_.grad.numpy()})
    np.savez(os.path.join(OUT, "gam_random.npz"), **cases)

    # ---- rp.gen_rp under torch.manual_seed, every distribution (rp.py:10-32)
    rpd = {}
    for dist in ["gaussian", "sphere", "very-sparse", "bernoulli", "uniform"]:
        for (d, k) in [(10, 1), (20, 1), (90, 5), (7, 3)]:
            torch.manual_seed(1234)
            rpd["%s_%d_%d" % (dist, d, k)] = rp.gen_rp(d, k, dist).numpy()
    # the J-projection weight matrix as training_routines.py:137,144-145 builds it
    torch.manual_seed(7)
    projs = [rp.gen_rp(10, 1, "gaussian") for _ in range(20)]
    rpd["weight_J20_d10"] = torch.cat(projs, dim=1).t().numpy()
    np.savez(os.path.join(OUT, "gen_rp.npz"), **rpd)

    # ---- rp.space_equally: Gram-Schmidt branch (d >= J, numpy RNG) and gradient-descent branch (d < J)
    se = {}
    np.random.seed(42)
    torch.manual_seed(42)
    P0 = torch.randn(6, 10)
    newW, _ = rp.space_equally(P0.clone(), lr=0.1, niter=5000)
    se["gs_in"] = P0.numpy()
    se["gs_out"] = newW.detach().numpy()
    np.random.seed(43)
    torch.manual_seed(43)
    P1 = torch.randn(8, 4)
    newW1, loss1 = rp.space_equally(P1.clone(), lr=0.1, niter=300)
    se["gd_in"] = P1.numpy()
    se["gd_out"] = newW1.detach().numpy()
    se["gd_loss"] = loss1.detach().numpy()
    np.savez(os.path.join(OUT, "space_equally.npz"), **se)
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    sys.exit(main_imq() if sys.argv[1:] == ["imq"] else main_synth() if sys.argv[1:] == ["synth"] else main())

"""GPU, world_size 2, NCCL: the row-partitioned operator (rows of K split over ranks, all-gather per product, all-gather
of dZ^ + all-reduce of the outputscale partials in the backward) gives the same MLL, gradients and predictions as the
single-process operator.  Skipped on boxes with fewer than two GPUs."""
import os
import socket
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    import training_routines as tr
    from rpgp import dist as rdist
    from rpgp import gp as gpytorch
    from rpgp.gp import settings
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        def run(partitioned):
            rdist.set_enabled(partitioned)
            torch.manual_seed(0)
            np.random.seed(0)
            g = torch.Generator().manual_seed(5)
            X = (torch.rand(1501, 6, generator=g) * 4 - 2).to(dev)            # uneven split: 751 + 750 rows
            y = (torch.sin(X).sum(-1) + 0.05 * torch.randn(1501, generator=g).to(dev))
            Xt = (torch.rand(300, 6, generator=g) * 4 - 2).to(dev)
            spec = tr.load_model_spec("additive_rp_prescale_J20")
            model, lik = tr.create_exact_gp(X, y, spec["kind"], **spec["model_kwargs"])
            model = model.to(dev)
            mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
            model.train()
            fixed = torch.randn(1501, 10, generator=torch.Generator().manual_seed(9)).to(dev)
            with settings.cg_tolerance(1e-4), settings.max_cg_iterations(2000), settings.deterministic_probes(fixed), \
                    settings.max_preconditioner_size(0), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                loss = -mll(model(X), y)
                loss.backward()
                model.eval()
                with torch.no_grad(), settings.eval_cg_tolerance(1e-5):
                    mean = model(Xt).mean
            grads = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}
            return loss.item(), grads, mean.cpu().numpy()

        loss_p, grads_p, mean_p = run(True)
        loss_s, grads_s, mean_s = run(False)
        assert abs(loss_p - loss_s) / abs(loss_s) < 1e-5, (loss_p, loss_s)
        for name in grads_s:
            a, b = grads_p[name], grads_s[name]
            # the symmetric tensor-core products accumulate with atomics (not bit-reproducible), so the two CG solves agree
            # to the CG tolerance (1e-4), not bit for bit; the mean gradient is a near-cancelling sum and gets a floor
            assert np.abs(a - b).max() <= 5e-3 * max(np.abs(b).max(), 1e-2), (name, a, b)
        assert np.abs(mean_p - mean_s).max() < 1e-4
        # every rank holds the same replicated result
        t = torch.tensor([loss_p], device=dev, dtype=torch.float64)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        assert all(float(x) == float(gathered[0]) for x in gathered)
        if rank == 0:
            np.save(out_path, np.array([loss_p, loss_s]))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_model_matches_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "loss.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    lp, ls = np.load(out)
    assert np.isfinite(lp) and abs(lp - ls) / abs(ls) < 1e-5

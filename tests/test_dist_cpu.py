"""CPU, world_size 2, gloo: the row partition and the one exchange of the K.V path (all-gather of row blocks per CG
iteration, all-reduce of the outputscale partials), and that CG over a row-partitioned product is identical on every
rank and equal to the single-process solve."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rpgp import dist as rdist
from rpgp.solver import linear_cg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)                      # replicated inputs: same seed on every rank
        A = torch.randn(n, n, dtype=torch.float64, generator=g)
        K = A @ A.t() / n + 2.0 * torch.eye(n, dtype=torch.float64)
        rhs = torch.randn(n, 3, dtype=torch.float64, generator=g)
        part = rdist.partition(n)
        assert part.world == world and part.rank == rank
        assert part.r0 == min(n, rank * part.block) and part.r1 == min(n, part.r0 + part.block)

        def matmul(V):                                            # each rank multiplies only its rows, then all-gather
            blk = K[part.r0:part.r1] @ V
            return rdist.all_gather_rows(blk, part)

        full = matmul(rhs)
        assert torch.allclose(full, K @ rhs, atol=1e-12)
        x = linear_cg(matmul, rhs, tolerance=1e-10, max_iter=500)
        partial = torch.tensor([float(rank + 1), 2.0 * (rank + 1)], dtype=torch.float64)
        total = rdist.all_reduce_sum(partial.clone())
        assert torch.allclose(total, torch.tensor([3.0, 6.0], dtype=torch.float64))
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        assert all(torch.equal(gathered[0], t) for t in gathered)  # bit-identical CG state on every rank
        if rank == 0:
            np.save(out_path, x.numpy())
        rdist.set_enabled(False)
        assert rdist.world_size() == 1 and rdist.partition(n).r1 == n
        rdist.set_enabled(True)
    finally:
        dist.destroy_process_group()


def test_row_partition_allgather_and_cg_world2(tmp_path):
    n = 101                                                       # uneven: blocks of 51 and 50 rows
    out = str(tmp_path / "x.npy")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    g = torch.Generator().manual_seed(0)
    A = torch.randn(n, n, dtype=torch.float64, generator=g)
    K = A @ A.t() / n + 2.0 * torch.eye(n, dtype=torch.float64)
    rhs = torch.randn(n, 3, dtype=torch.float64, generator=g)
    x_single = linear_cg(K.matmul, rhs, tolerance=1e-10, max_iter=500)
    np.testing.assert_allclose(np.load(out), x_single.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(np.load(out), torch.linalg.solve(K, rhs).numpy(), atol=1e-7)


def _rect_worker(rank, world, port, m, n):
    """the row-partitioned rectangular product (prediction: test rows x training columns) with the fused kernel replaced by an
    explicit matrix: every rank multiplies only its block of test rows, one all-gather rebuilds the product"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rpgp import lazy, ops
        g = torch.Generator().manual_seed(1)
        Kx = torch.randn(m, n, dtype=torch.float64, generator=g)
        V = torch.randn(n, 2, dtype=torch.float64, generator=g)
        calls = []

        def fake_kmv_raw(Z1, Z2, c, J, K, V_, packed1=None, packed2=None, nlc=None, row_range=None, base=0):
            calls.append(row_range)
            r0, r1 = row_range if row_range is not None else (0, Z1.shape[0])
            return Kx[r0:r1] @ V_

        real = ops.kmv_raw
        ops.kmv_raw = fake_kmv_raw
        try:
            Z1, Z2 = torch.zeros(m, 3, dtype=torch.float64), torch.zeros(n, 3, dtype=torch.float64)
            out = ops.kmv_rect_partitioned(Z1, Z2, torch.ones(3, dtype=torch.float64), 3, 1, V)
            assert torch.allclose(out, Kx @ V, atol=1e-12) and out.shape == (m, 2)
            part = rdist.partition(m)
            assert calls == [(part.r0, part.r1)] and part.r1 - part.r0 < m
            # the operator takes that route for rectangular products when a process group exists
            op = lazy.RPAdditiveLazyTensor(Z1, Z2, torch.ones(3, dtype=torch.float64), 3, 1)
            assert torch.allclose(op._matmul(V), Kx @ V, atol=1e-12) and calls[-1] == (part.r0, part.r1)
            # too few rows per rank: replicated
            small = ops.kmv_rect_partitioned(Z1[:10], Z2, torch.ones(3, dtype=torch.float64), 3, 1, V)
            assert calls[-1] is None and torch.allclose(small, Kx[:10] @ V, atol=1e-12)
        finally:
            ops.kmv_raw = real
    finally:
        dist.destroy_process_group()


def test_rectangular_product_row_partition_world2():
    mp.spawn(_rect_worker, args=(2, _free_port(), 301, 77), nprocs=2, join=True)      # 151 + 150 test rows


def _probe_worker(rank, world, port, out_dir):
    """ranks with DIFFERENT local RNG streams: the SLQ probes are drawn on rank 0 and broadcast, so the stochastic
    log-determinant is identical on every rank; operands that differ between ranks are reported, not silently reduced"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rpgp.gp import settings
        from rpgp.lazy import AddedDiagLazyTensor, DenseLazyTensor
        g = torch.Generator().manual_seed(3)
        A = torch.randn(60, 60, dtype=torch.float64, generator=g)
        Kmat = (A @ A.t() / 60).requires_grad_(True)
        y = torch.randn(60, 1, dtype=torch.float64, generator=g)
        noise = torch.tensor(0.5, dtype=torch.float64, requires_grad=True)
        torch.manual_seed(100 + rank)                              # the local streams differ
        with settings.max_cholesky_size(0), settings.cg_tolerance(1e-10), settings.max_cg_iterations(500):
            iq, ld = AddedDiagLazyTensor(DenseLazyTensor(Kmat), noise).inv_quad_logdet(inv_quad_rhs=y, logdet=True)
            (iq + ld).backward()
        vals = torch.stack([iq.detach(), ld.detach(), noise.grad])
        gathered = [torch.empty_like(vals) for _ in range(world)]
        dist.all_gather(gathered, vals)
        assert all(torch.equal(gathered[0], v) for v in gathered), gathered
        rdist.assert_replicated("identical tensors", Kmat, y)
        try:
            rdist.assert_replicated("a rank-dependent tensor", y + rank)
            raised = False
        except RuntimeError as e:
            raised = "differs between rank 0 and rank 1" in str(e)
        assert raised
        x = torch.full((4,), float(rank))
        assert torch.equal(rdist.broadcast_(x), torch.zeros(4))
    finally:
        dist.destroy_process_group()


def test_slq_probes_are_broadcast_and_replication_is_checked(tmp_path):
    mp.spawn(_probe_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)


def test_partition_arithmetic():
    p = rdist.Partition(10, 4, 3)
    assert (p.block, p.r0, p.r1) == (3, 9, 10)
    assert [rdist.Partition(10, 4, r).rows() for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert rdist.Partition(2, 4, 3).rows() == (2, 2)              # more ranks than rows: empty block
    assert rdist.partition(7).world == 1

"""Row partition of K over the ranks of one 8xB200 box, and the one exchange the path has.

One process per GPU (torchrun); rank r owns the contiguous rows [r*ceil(n/G), (r+1)*ceil(n/G)) of K.  X, W, the
hyper-parameters and the CG state are replicated, so Z^ is computed redundantly on every rank (no exchange) and every
CG iteration needs exactly one collective: an all-gather of the (n/G x t) row blocks of K.P (NCCL over NVLink on GPUs;
gloo in the CPU tests) -- or, with the symmetric tensor-core kernel, an all-reduce of the partial products.
REPLICATION CONTRACT: every rank must hold bit-identical X, y, hyper-parameters and right-hand sides.  Random draws made
inside the solver (the SLQ probe vectors) are drawn on rank 0 and broadcast (`broadcast_`); `assert_replicated` checks the
operands of every MLL solve and raises when ranks disagree (differently seeded ranks used to give silently wrong results).  Gradients: the rows' dZ^ are all-gathered once per step, the J outputscale partials are
all-reduced.  This replaces gpytorch.kernels.MultiDeviceKernel (training_routines.py:407-408), which scatters x1 over
devices inside one process and copies row blocks back through peer memcpys.
"""
import torch
import torch.distributed as tdist


class Partition:
    __slots__ = ("n", "world", "rank", "block", "r0", "r1")

    def __init__(self, n, world, rank):
        self.n, self.world, self.rank = int(n), int(world), int(rank)
        self.block = (self.n + self.world - 1) // self.world
        self.r0 = min(self.n, self.rank * self.block)
        self.r1 = min(self.n, self.r0 + self.block)

    def rows(self, rank=None):
        r = self.rank if rank is None else rank
        r0 = min(self.n, r * self.block)
        return r0, min(self.n, r0 + self.block)


_enabled = True


def set_enabled(flag):
    """Turn the row partition off (every rank computes everything) -- used by single-rank tools inside a job."""
    global _enabled
    _enabled = bool(flag)


def world_size():
    if _enabled and tdist.is_available() and tdist.is_initialized():
        return tdist.get_world_size()
    return 1


def rank():
    if _enabled and tdist.is_available() and tdist.is_initialized():
        return tdist.get_rank()
    return 0


def partition(n):
    return Partition(n, world_size(), rank())


def all_gather_rows(block, part):
    """Row blocks (part.r1-part.r0 x t) of every rank -> the full (n x t) matrix on every rank."""
    if part.world == 1:
        return block
    t = block.shape[1:]
    if block.shape[0] < part.block:  # last rank(s) may own fewer rows: pad to the common block size
        pad = block.new_zeros((part.block - block.shape[0],) + tuple(t))
        block = torch.cat([block, pad], dim=0)
    full = block.new_empty((part.block * part.world,) + tuple(t))
    tdist.all_gather_into_tensor(full, block.contiguous())
    return full[:part.n]


def broadcast_(x, src=0):
    """in-place broadcast from rank `src` (random draws -- SLQ probes, initial values -- must be IDENTICAL on every rank: the CG
    state is replicated, and a rank multiplying different vectors would all-reduce garbage without any error)"""
    if world_size() > 1:
        tdist.broadcast(x, src=src)
    return x


def assert_replicated(what, *tensors, rtol=0.0):
    """Raise if a tensor that must be replicated differs between ranks (checksum = float64 sum and sum of squares, compared
    exactly by default).  One small all-gather; called once per solve on the right-hand sides and the operator's representation."""
    if world_size() == 1:
        return
    sums = []
    for t in tensors:
        td = t.detach().double()
        sums += [td.sum(), (td * td).sum()]
    mine = torch.stack(sums)
    allv = [torch.empty_like(mine) for _ in range(world_size())]
    tdist.all_gather(allv, mine)
    ref = allv[0]
    for r, v in enumerate(allv):
        bad = (v - ref).abs() > rtol * ref.abs()
        if bool(bad.any()):
            raise RuntimeError("rpgp.dist: %s differs between rank 0 and rank %d (checksums %s vs %s): every rank must hold identical "
                               "replicated state -- seed all ranks identically or broadcast the random draws" %
                               (what, r, ref.tolist(), v.tolist()))


def all_reduce_sum(x):
    if world_size() > 1:
        tdist.all_reduce(x, op=tdist.ReduceOp.SUM)
    return x

#!/usr/bin/env python
"""bench.py -- throughput of the K.V hot path (BASELINE.json metric: kernel-MVM pair-evals/s and CG-MLL iters/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one CG iteration of the exact-GP MLL solve at the named workload: one K^.P product with P in R^{n x t}
(the fused sm_100a kernel, rows of K partitioned over the N ranks, NCCL all-gather of the row blocks) plus the O(n t)
CG vector updates.  `value` = (i, i', j) pair-evaluations per second for the whole job (n*n*J per step), inputs
resident in HBM.  `e2e` = the same metric through the host-buffer C-ABI entry point (rpgp_kmv_host_f32: H2D of X and V
from pinned memory, projection, K.V, D2H of the product) every step.  `roofline` is the MUFU (XU-pipe ex2) roof the
north star names, with the denominator measured live by the library's own microbenchmark (rpgp_measure_peaks); the
HBM view (algorithmic bytes vs MEASURED_PEAKS.json) is reported beside it.  `cpu_baseline` / `--impl reference` time
the C restatement of the reference's dense arithmetic (oracle/kmv_oracle.c) on the host cores, on a bounded row sample.
`parity` compares sampled rows of the LAST timed product (after the all-reduce when N > 1) with the FP64 C oracle applied
to the same packed coordinates and the same right-hand sides; above the tolerance the run exits non-zero.  `mll_step` is the
north-star quantity itself -- one exact MLL + gradient evaluation through the reference-facing model API -- at cfg2.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "randomly-projected-additive-gps_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CPU_BUILD = "oracle/kmv_oracle.c, gcc -O3 -mavx2 -mfma -ffast-math -pthread, float32 (C port: the reference is pure Python over GPyTorch/KeOps)"

WORKLOADS = {
    # name: n, d, J, K, t, projection, c, description (BASELINE.json configs[i])
    "cfg1": dict(n=2_000, d=10, J=20, K=1, t=11, proj="gaussian", desc="configs[0] n=2k d=10 additive_rp_J20_K1"),
    "cfg2": dict(n=100_000, d=20, J=20, K=1, t=11, proj="gaussian", desc="configs[1] n=100k d=20 additive_rp_prescale_J20"),
    "cfg3": dict(n=400_000, d=26, J=26, K=1, t=16, proj="spread", desc="configs[2] n=400k d=26 DPA-GP additive_spread_prescale_Jd t=16"),
    "cfg4": dict(n=1_000_000, d=90, J=20, K=1, t=11, proj="gaussian", desc="configs[3] n=1M d=90 J=20 K=1 songs-shaped"),
    "cfg5a": dict(n=1_000_000, d=90, J=1, K=20, t=11, proj="gaussian", desc="configs[4] n=1M d=90 J=1 K=20"),
    "cfg5b": dict(n=1_000_000, d=90, J=20, K=5, t=11, proj="gaussian", desc="configs[4] n=1M d=90 J=20 K=5"),
}


def make_inputs(w, seed=0):
    """Synthetic inputs of SURVEY.md §8(d): X ~ N(0,1), W from gen_rp (or diversified rows), ell = 1, c = ln2/J,
    sigma_n^2 = 1, probes V column-normalised.  Host (pinned when CUDA is present) float32 tensors."""
    import rp
    torch.manual_seed(seed)
    np.random.seed(seed)
    n, d, J, K, t = w["n"], w["d"], w["J"], w["K"], w["t"]
    X = torch.randn(n, d)
    projs = [rp.gen_rp(d, K, "gaussian") for _ in range(J)]
    W = torch.cat(projs, dim=1).t().contiguous()
    if w["proj"] == "spread":
        W, _ = rp.space_equally(W, lr=0.1, niter=5000)
        W = W.contiguous()
        c = torch.full((J,), math.log(2.0))                       # DPA-GP: no 1/J (training_routines.py:168)
        inv_ell = torch.full((d,), 1.0 / math.log(2.0)) * 1.0     # base lengthscale softplus(0) folded into ell
    else:
        c = torch.full((J,), math.log(2.0) / J)
        inv_ell = torch.ones(d)
    V = torch.randn(n, t)
    V = V / V.norm(dim=0, keepdim=True)
    if torch.cuda.is_available():
        X, V = X.pin_memory(), V.pin_memory()
    return X, W, inv_ell, c, V


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([f.strip() for f in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, flag in zip(names, r[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def cpu_sample_rate(Z, c, J, K, V, n, seconds=12.0, threads=0):
    """Time the C oracle on a bounded row sample; returns (pair-evals/s, rows used, threads, seconds)."""
    from oracle import c_oracle
    threads = threads or c_oracle.max_threads()
    quantum = 16 * threads                                      # kmv_oracle.c hands out 16-row blocks round-robin
    probe_rows = min(n, 2 * quantum)
    t0 = time.perf_counter()
    c_oracle.kmv(Z[:probe_rows], Z, c, J, K, V, threads=threads)
    dt = time.perf_counter() - t0
    rate = probe_rows * n * J / dt
    rows = int(min(n, max(probe_rows, rate * seconds / (n * J))))
    rows = min(n, max(quantum, (rows // quantum) * quantum))
    t0 = time.perf_counter()
    c_oracle.kmv(Z[:rows], Z, c, J, K, V, threads=threads)
    dt = time.perf_counter() - t0
    return rows * n * J / dt, rows, threads, dt


def sample_rows(n, count=256, seed=1):
    """row indices for the at-scale parity check: the first and the last 32 rows (the last 128-row block is partial unless
    128 | n) plus runs of 8 consecutive rows at random positions, so that every role of a row (row side of its own block pairs,
    column side of the others', any rank's share) is covered"""
    rng = np.random.RandomState(seed)
    rows = set(range(min(32, n))) | set(range(max(0, n - 32), n))
    while len(rows) < min(count, n):
        r0 = int(rng.randint(0, max(1, n - 8)))
        rows.update(range(r0, min(n, r0 + 8)))
    return np.array(sorted(rows)[:count], dtype=np.int64)


def natural_f64(zp, lay, J, K):
    """packed, pre-scaled planes (nchunks, n, CP) float32 -> natural (n, J*K) float64 coordinates holding EXACTLY the values the
    kernels see (the oracle's exp(-d^2/2) of z/scale equals the kernels' 2^(-|dz|^2))"""
    from rpgp import _lib
    z = zp.cpu().numpy()
    nch, n, CP = z.shape
    g = z[:, :, :lay.G * lay.KP].reshape(nch, n, lay.G, lay.KP)[..., :K]
    g = np.transpose(g, (1, 0, 2, 3)).reshape(n, nch * lay.G, K)[:, :J, :]
    return np.ascontiguousarray(g.reshape(n, J * K), dtype=np.float64) / _lib.coord_scale()


def parity_record(zp, lay, c, J, K, P, KP_gpu, tol=1e-5, count=256):
    """sampled rows of the GPU product K.P against oracle_kmv_f64 (oracle/kmv_oracle.c) on identical inputs"""
    from oracle import c_oracle
    n = zp.shape[1]
    rows = sample_rows(n, count)
    Zn = natural_f64(zp, lay, J, K)
    t0 = time.perf_counter()
    ref = c_oracle.kmv(Zn[rows], Zn, np.asarray(c, np.float64), J, K, P.double().cpu().numpy(), dtype=np.float64)
    secs = time.perf_counter() - t0
    got = KP_gpu[torch.as_tensor(rows, device=KP_gpu.device)].double().cpu().numpy()
    diff = got - ref
    row_rel = np.linalg.norm(diff, axis=1) / np.maximum(np.linalg.norm(ref, axis=1), 1e-300)
    rec = {"rows": int(len(rows)), "columns": int(n), "norm_rel": float(np.linalg.norm(diff) / np.linalg.norm(ref)),
           "max_row_rel": float(row_rel.max()), "max_abs": float(np.abs(diff).max()), "ref_max_abs": float(np.abs(ref).max()),
           "tol": tol, "oracle": "oracle_kmv_f64 (oracle/kmv_oracle.c) on the packed FP32 coordinates and the right-hand sides of the "
                                 "last timed product", "oracle_seconds": secs}
    rec["ok"] = bool(rec["norm_rel"] <= tol and rec["max_row_rel"] <= tol)
    return rec


def run_reference(args, w):
    """--impl reference: the reference's CPU arithmetic for the path (C port of the oracle: the reference is pure Python
    over GPyTorch/KeOps and cannot be installed or compiled here -- DESIGN.md), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle, rpgp_oracle as orc
    X, W, inv_ell, c, V = make_inputs(w)
    n, J, K = w["n"], w["J"], w["K"]
    Z = orc.scaled_projection(X.numpy(), W.numpy(), 1.0 / inv_ell.numpy(), prescale=True, dtype=np.float32)
    threads = c_oracle.max_threads()
    quantum = 16 * threads                                      # kmv_oracle.c hands out 16-row blocks round-robin
    probe_rows = min(n, 2 * quantum)
    t0 = time.perf_counter()
    c_oracle.kmv(Z[:probe_rows], Z, c.numpy(), J, K, V.numpy(), threads=threads)
    rate = probe_rows * n * J / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)              # whole run within a few minutes
    rows = int(min(n, max(quantum, rate * min(budget, 8.0) / (n * J))))
    rows = min(n, max(quantum, (rows // quantum) * quantum))
    for _ in range(args.warmup):
        c_oracle.kmv(Z[:rows], Z, c.numpy(), J, K, V.numpy(), threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.kmv(Z[:rows], Z, c.numpy(), J, K, V.numpy(), threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = rows * n * J / dt
    sample = "%d of %d rows x all %d columns per step (extrapolates linearly in rows)" % (rows, n, n)
    line = {
        "impl": "reference", "metric": "kernel_mvm_pair_evals_per_s", "value": value, "unit": "pair-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "ms_per_full_step_extrapolated": dt * 1e3 * n / rows,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "n": n, "d": w["d"], "J": J, "K": K, "t": w["t"]},
        "cpu_baseline": {"value": value, "unit": "pair-evals/s", "cores": threads, "kind": "port", "sample": sample,
                         "build": CPU_BUILD},
        "e2e": {"value": value, "unit": "pair-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cg_iters_per_s": 1.0 / (dt * n / rows), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, w):
    from rpgp import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; the K.V path has no CPU fallback")
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    import torch.distributed as dist

    n, d, J, K, t = w["n"], w["d"], w["J"], w["K"], w["t"]
    X, W, inv_ell, c, V = make_inputs(w)
    lay = _lib.plan_layout(J, K)
    noise = 1.0

    # row partition: rank r owns rows [r*blk, min(n, (r+1)*blk))
    blk = (n + world - 1) // world
    r0, r1 = min(n, rank * blk), min(n, (rank + 1) * blk)

    Xd, Wd, Vd = X.to(dev, non_blocking=True), W.to(dev), V.to(dev, non_blocking=True)
    zp = _lib.project(Xd, Wd, inv_ell.to(dev), None, lay)       # Z^ computed once per MLL step, before the CG loop
    del Xd
    nlc = _lib.pack_log2c(c.to(dev), lay)

    use_sym = (not args.no_sym) and _lib.mvm_sym_supported(lay, t)
    nblocks128 = (n + 127) // 128
    sper = (nblocks128 + world - 1) // world
    sb0, sb1 = min(nblocks128, rank * sper), min(nblocks128, (rank + 1) * sper)

    # CG state (replicated on every rank, updated identically): solve K^ x = V
    x = torch.zeros_like(Vd)
    r = Vd.clone()
    p = r.clone()
    rz = (r * r).sum(0)
    Kp_full = torch.empty((blk * world, t), device=dev)
    kernel_events = []
    last = {}

    def cg_iteration(record):
        nonlocal rz, p
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev = (e0, e1)
            kernel_events.append(ev)
        else:
            ev = None
        if use_sym:
            # symmetric tensor-core kernel: this rank's share of the unique 128-row block pairs, partial sums for all rows
            Kp = _lib.mvm_sym(zp, lay, nlc, p, block_range=(sb0, sb1), events=ev)
            if world > 1:
                dist.all_reduce(Kp, op=dist.ReduceOp.SUM)
        else:
            Kp_blk = _lib.mvm_fwd(zp, zp, lay, nlc, p, row_range=(r0, r1), events=ev)
            if world > 1:
                if Kp_blk.shape[0] < blk:
                    Kp_blk = torch.cat([Kp_blk, Kp_blk.new_zeros((blk - Kp_blk.shape[0], t))])
                dist.all_gather_into_tensor(Kp_full, Kp_blk.contiguous())
                Kp = Kp_full[:n].clone()
            else:
                Kp = Kp_blk
        last["p"], last["Kp"] = p, Kp            # (neither is modified in place below) operands of the at-scale parity check
        Kp = Kp + noise * p
        alpha = rz / (p * Kp).sum(0).clamp_min(1e-30)
        x.add_(p * alpha)
        r.sub_(Kp * alpha)
        rz_new = (r * r).sum(0)
        p = r + p * (rz_new / rz.clamp_min(1e-30))
        rz = rz_new

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        cg_iteration(False)
    sync()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record()
        for _ in range(args.steps):
            cg_iteration(True)
        e1.record()
        sync()
    launches = _lib.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kernel_events) / max(1, len(kernel_events))], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    kernel_ms = float(kms.item())
    value = float(n) * n * J / (ms_per_step * 1e-3)

    # ---- parity at the benchmarked scale: sampled rows of the last timed product vs the FP64 C oracle (rank 0; the product is
    # replicated after the collective) --------------------------------------------------------------------------------------
    parity = None
    if not args.no_parity and rank == 0:
        parity = parity_record(zp, lay, c.numpy(), J, K, last["p"], last["Kp"], count=args.parity_rows)
        parity["after_collective"] = world > 1
    last.clear()

    # ---- end-to-end through the host-buffer C ABI: every step H2D of X, W, scales, c and V from pinned memory, projection,
    # the SAME symmetric product on this rank's block pairs, the all-reduce over the ranks, sigma^2 V, D2H of the rank's rows ------
    e2e = None
    if not args.no_e2e:
        Xn, Wn, Vn = X.numpy(), W.numpy(), V.numpy()
        ie, cn = inv_ell.numpy(), c.numpy()
        plan = _lib.HostPlan(n, d, J, K, t, device=local)
        out_host = torch.empty((r1 - r0, t), dtype=torch.float32).pin_memory().numpy()
        reduce_fn = (lambda buf: dist.all_reduce(buf, op=dist.ReduceOp.SUM)) if world > 1 else None

        def host_step():
            plan.set_operator(Xn, Wn, ie, None, cn)
            return plan.kmv(Vn, diag_add=noise, block_range=(sb0, sb1) if use_sym else (r0 // 128, (r1 + 127) // 128),
                            row_range=(r0, r1), all_reduce=reduce_fn, out=out_host)
        for _ in range(2):
            host_step()
        sync()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()
        sync()
        dt = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], device=dev, dtype=torch.float64)
        e2e_launches = (_lib.launch_count() - l0) // args.e2e_steps
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = 4 * (n * d + J * K * d + d + J + n * t)
        e2e = {"value": float(n) * n * J / float(dt.item()), "unit": "pair-evals/s", "ms_per_step": float(dt.item()) * 1e3,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(4 * (r1 - r0) * t), "gpu_launches_per_step": int(e2e_launches),
               "vs_device_timed": float(dt.item()) * 1e3 / ms_per_step,
               "api": "rpgp_plan_set_operator + rpgp_plan_kmv_begin [+ NCCL all-reduce of the partial products] + rpgp_plan_kmv_end "
                      "(host buffers, pinned; projection + the same symmetric K.V + sigma^2 V; device buffers owned by the plan)"}
        if world == 1 and use_sym and not args.no_parity:
            # the host path's rows against the device path's on the same operands
            ref_rows = (_lib.mvm_sym(zp, lay, nlc, Vd) + noise * Vd)[r0:r1].cpu().numpy()
            e2e["vs_device_path_rel"] = float(np.linalg.norm(out_host - ref_rows) / np.linalg.norm(ref_rows))
        plan.close()

    mll = None
    if args.mll_workload != "none":
        del x, r, p, Kp_full, zp
        _lib.free_workspaces()
        torch.cuda.empty_cache()
        mll = mll_step_record(args, WORKLOADS[args.mll_workload], dev, world, rank, steps=args.mll_steps, warmup=1)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline: MUFU roof measured live by the library's own microbenchmark -------------------------------------------
    peaks = _lib.measure_peaks()
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    mufu_peak = peaks["mufu_ex2"]["mufu_per_clk_sm"] * sms * peaks["mufu_ex2"]["mhz"] * 1e6
    fp32_peak = peaks["ffma2"]["fp32_per_clk_sm"] * sms * peaks["ffma2"]["mhz"] * 1e6
    m_rows = r1 - r0  # rank 0's block (largest)
    groups = lay.nchunks * lay.G  # ex2 actually issued per pair (padding groups included)
    ex2_alg = float(m_rows) * n * J
    ach = ex2_alg / (kernel_ms * 1e-3)
    TPb = _lib.padded_rhs(lay, min(t, 16), False)
    # exponential pairs the kernels hand to the FMA-pipe polynomial (csrc/dispatch.cuh default_poly_pairs, sym_tc5.cu launcher)
    if K != 1:
        np2 = 0
    elif use_sym:
        np2 = int(os.environ.get("RPGP_SYM_POLY_PAIRS", 2 if lay.CP >= 20 else (1 if lay.CP >= 16 else 0)))
    else:
        np2 = int(os.environ.get("RPGP_POLY_PAIRS", 0 if TPb > 16 else (3 if lay.CP >= 28 else 2 if lay.CP >= 16 else 1 if lay.CP >= 8 else 0)))
    evals = (float(n) * n / 2 / world) if use_sym else float(m_rows) * n          # kernel values actually formed by this rank
    # FP32-pipe lane-operations the kernel EXECUTES per kernel value (instruction mix of csrc/kv_kernels.cuh pair_kernel_value and
    # the S split of sym_tc5.cu; a packed FADD2 / FFMA2 counts two): K = 1: sub, fma, add per coordinate + 18 per polynomial
    # exponential; K > 1 direct differences: 2 per coordinate + 1 add per group; symmetric kernels add the tf32 split (1 sub per value)
    # and spend nothing on V (tensor cores); the SIMT kernel spends t fma per value on V
    if K == 1:
        fp32_exec_per_value = 3.0 * lay.nchunks * lay.CP + 18.0 * 2 * np2
    else:
        fp32_exec_per_value = lay.nchunks * lay.G * (2.0 * lay.KP + 1.0)
    fp32_exec_per_value += 1.0 if use_sym else float(TPb)
    # K > 1 symmetric products: squared distances on tcgen05 (csrc/sym_tcd.cu) while the centred coordinates stay inside the gate
    tcd = _lib.mvm_sym_distance_plan(lay) if (use_sym and K > 1) else None
    if tcd is not None:
        zp = _lib.project(X.to(dev), W.to(dev), inv_ell.to(dev), None, lay)
        zc = zp - zp.mean(dim=1, keepdim=True)                 # packed planes are already scaled by sqrt(log2(e)/2)
        r2 = torch.stack([(zc[ch, :, g * lay.KP:g * lay.KP + K] ** 2).sum(-1)
                          for ch in range(lay.nchunks) for g in range(lay.G) if ch * lay.G + g < J])
        max_norm2, rms_norm2 = float(r2.max()), float((r2.double() ** 2).mean().sqrt())
        del zc, r2, zp
        tcd["max_centred_norm2"], tcd["rms_centred_norm2"] = max_norm2, rms_norm2
        if rms_norm2 > tcd["bound"] or max_norm2 > 10 * tcd["bound"]:      # the device-side gate of csrc/sym_tcd.cu (tcd_gate)
            tcd = dict(tcd, active=False)
        else:
            tcd = dict(tcd, active=True)
            groups = tcd["nchunks"] * tcd["groups_per_chunk"]
            fp32_exec_per_value = groups * 1.0 + 1.0            # one add per group (the exponent comes from TMEM) + the split
    xu_ex2 = evals * (groups - 2 * np2)
    xu_frac = xu_ex2 / (kernel_ms * 1e-3) / mufu_peak
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    alg_bytes = 4.0 * (m_rows * lay.nchunks * lay.CP + n * lay.nchunks * lay.CP + n * _lib.padded_rhs(lay, min(t, 16), False) + m_rows * t)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
    except Exception:
        pass
    roofline = {
        "bound": "mufu", "achieved": xu_ex2 / (kernel_ms * 1e-3) / 1e12, "peak": mufu_peak / 1e12, "unit": "Tex2/s", "frac": xu_frac,
        "frac_definition": "XU (MUFU) pipe utilisation: MUFU.EX2 the kernel actually issues per launch / kernel time / measured peak",
        "x_nonsym_roof": ach / mufu_peak,
        "x_nonsym_roof_definition": "ALGORITHMIC exponentials of SURVEY 8(d) (m*n*J per product) / kernel time / measured peak; exceeds 1 "
                                    "because " + ("symmetry halves the evaluations and " if use_sym else "")
                                    + "%d of every %d exponentials are evaluated by an FMA-pipe polynomial" % (2 * np2, groups),
        "peak_source": "measured live: rpgp_measure_peaks mufu_ex2 %.2f/clk/SM x %d SMs x %.0f MHz"
                       % (peaks["mufu_ex2"]["mufu_per_clk_sm"], sms, peaks["mufu_ex2"]["mhz"]),
        "kernel": ("mvm_sym_tcd_kernel<NL=%d> (symmetric; squared distances as augmented inner products on tcgen05 kind::tf32, "
                   "3xTF32, %d groups x %d k-steps per chunk; S.V and S^T.V on tcgen05 as well)"
                   % (tcd["lines"], tcd["groups_per_chunk"], tcd["ksteps_per_group"])) if (tcd is not None and tcd["active"])
                  else ("mvm_sym_tc5_kernel<CP=%d,NP2=%d> (symmetric: each kernel value evaluated once; S.V and S^T.V both on tcgen05 "
                        "kind::tf32, 3xTF32 split)" % (lay.CP, np2)) if use_sym
                  else "mvm_fwd_kernel<CP=%d,TP=%d,KP=%d,G=%d,NP2=%d>" % (lay.CP, TPb, lay.KP, lay.G, np2),
        "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / ms_per_step,
        "algorithmic_ex2_per_launch": ex2_alg, "xu_ex2_per_launch": xu_ex2,
        "fp32_frac": evals * fp32_exec_per_value / (kernel_ms * 1e-3) / fp32_peak,
        "fp32_frac_definition": "FP32-pipe lane-operations the kernel executes (modelled from its instruction mix: %.0f per kernel value) "
                                "/ kernel time / measured FFMA2 peak" % fp32_exec_per_value,
        "fp32_peak_Tlaneops": fp32_peak / 1e12,
        "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one ncu capture of "
                                              "this kernel at this workload, per launch)" if traffic is not None else None,
        "distance_on_tensor_cores": tcd,
        "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_GBps": alg_bytes / (kernel_ms * 1e-3) / 1e9,
                "peak_GBps": hbm_peak, "frac": (alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None,
                "peak_source": "MEASURED_PEAKS.json" if hbm_peak else "absent"},
    }

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import rpgp_oracle as orc
        Z = orc.scaled_projection(X.numpy(), W.numpy(), 1.0 / inv_ell.numpy(), prescale=True, dtype=np.float32)
        rate, rows, threads, secs = cpu_sample_rate(Z, c.numpy(), J, K, V.numpy(), n)
        cpu = {"value": rate, "unit": "pair-evals/s", "cores": threads, "kind": "port", "build": CPU_BUILD,
               "sample": "%d of %d rows x all %d columns, %.1f s" % (rows, n, n, secs)}

    line = {
        "metric": "kernel_mvm_pair_evals_per_s", "value": value, "unit": "pair-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "n": n, "d": d, "J": J, "K": K, "t": t},
        "run": {"noise": noise,
                "parallelism": ("unique 128-row block pairs of the symmetric K over %d rank(s), NCCL all-reduce(sum) of the partial "
                                "products per CG iteration" if use_sym else
                                "rows of K over %d rank(s), NCCL all-gather per CG iteration") % world,
                "l2": "inputs larger than L2 (Z^ %.0f MB, V %.0f MB)" % (n * lay.nchunks * lay.CP * 4 / 1e6, n * t * 4 / 1e6)
                if n * lay.nchunks * lay.CP * 4 > 126e6 else "inputs fit in L2 (reused every step by design: Z^ is read n/256 times per launch)"},
        "cg_iters_per_s": 1e3 / ms_per_step, "pairs_per_s": value / J,
        "e2e": e2e, "gpu_launches": int(launches), "parity": parity, "roofline": roofline, "cpu_baseline": cpu, "mll_step": mll,
        "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: PARITY FAILURE %s\n" % json.dumps(parity))
        return 3
    return 0


def mll_step_record(args, w, dev, world, rank, steps=1, warmup=1):
    """One full training step through the reference-facing API -- model(X) -> -MLL -> backward (pivoted-Cholesky preconditioner,
    multi-RHS CG with SLQ probes, fused gradient kernel; fitting/optimizing.py:65-74) -- with the solver settings of the reference's
    large runs (run_scripts/additive_spread_prescale_Jd.sh:6: --cg_tol 0.002).  Returns the record (identical on every rank)."""
    import importlib
    import warnings

    import training_routines as tr
    from rpgp import _lib, gp as gpytorch
    cg_mod = importlib.import_module("rpgp.solver.linear_cg")   # the module (rpgp.solver re-exports the function under the same name)
    torch.manual_seed(0)
    np.random.seed(0)
    n, d, J, K = w["n"], w["d"], w["J"], w["K"]
    gen = torch.Generator(device="cpu").manual_seed(0)
    X = torch.randn(n, d, generator=gen).to(dev)
    wtrue = (torch.randn(d, 8, generator=gen) / math.sqrt(d)).to(dev)
    y = torch.sin(X @ wtrue).sum(-1) + 0.1 * torch.randn(n, generator=gen).to(dev)
    y = (y - y.mean()) / y.std()
    kw = dict(J=J, k=K, noise_prior=True, kernel_type="RBF", learn_proj=False, prescale=True, batch_kernel=(K == 1))
    if w["proj"] == "spread":
        kw.update(space_proj=True, batch_kernel=False, mem_efficient=True)
    model, lik = tr.create_exact_gp(X, y, "additive_rp", **kw)
    model = model.to(dev)
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    model.train()
    fwd_ev, bwd_ev = [], []

    def step(record):
        model.zero_grad()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        loss = -mll(model(X), y)
        ev[1].record()
        loss.backward()
        ev[2].record()
        if record:
            fwd_ev.append((ev[0], ev[1]))
            bwd_ev.append((ev[1], ev[2]))
        return loss

    max_it = args.mll_max_cg if args.mll_max_cg > 0 else 10_000
    with gpytorch.settings.cg_tolerance(args.cg_tol), gpytorch.settings.max_cg_iterations(max_it), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(warmup):
            step(False)
        torch.cuda.synchronize(dev)
        it0, l0 = cg_mod.STATS["iterations"], _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(dev.index or 0) as clocks, _lib.timing() as tm:
            e0.record()
            for _ in range(steps):
                loss = step(True)
            e1.record()
            torch.cuda.synchronize(dev)
            split = tm.totals()
    ms = e0.elapsed_time(e1) / steps
    iters = (cg_mod.STATS["iterations"] - it0) / steps
    fwd_ms = sum(a.elapsed_time(b) for a, b in fwd_ev) / steps
    bwd_ms = sum(a.elapsed_time(b) for a, b in bwd_ev) / steps
    kv_calls, kv_ms = 0, 0.0
    for name in ("mvm_sym", "mvm_fwd"):
        if name in split:
            kv_calls += split[name][0]
            kv_ms += split[name][1]
    grad_calls, grad_ms = split.get("quad_bwd", (0, 0.0))
    kv_calls, kv_ms, grad_ms = kv_calls / steps, kv_ms / steps, grad_ms / steps
    return {"metric": "mll_grad_step", "ms_per_step": ms, "steps_per_s": 1e3 / ms, "cg_iterations_per_step": iters,
            "cg_iters_per_s": iters / (ms * 1e-3), "pair_evals_per_s": (kv_calls + 2.6) * float(n) * n * J / (ms * 1e-3),
            "pair_evals_note": "K.V products of the step + the gradient pass counted as 2.6 products (it evaluates every pair twice "
                               "with 5J+2t FP32 lane-ops), x n^2 J, / step time",
            "loss": float(loss), "n_gpus": world, "steps": steps, "warmup": warmup, "cg_tol": args.cg_tol,
            "max_cg_iterations": max_it, "gpu_launches": (_lib.launch_count() - l0) // steps,
            "split_ms": {"forward_total": fwd_ms, "backward_total": bwd_ms, "kv_products": kv_ms, "kv_product_calls": kv_calls,
                         "gradient_kernel": grad_ms, "solver_and_host_overhead": ms - kv_ms - grad_ms},
            "clocks": clocks.summary(),
            "config": {"workload": w["desc"], "n": n, "d": d, "J": J, "K": K, "t": 11}}


def run_mll(args, w):
    """--mode mll: only the full MLL + gradient step at the named workload (auxiliary line; the headline is the default mode,
    which carries the same record at cfg2 under `mll_step`)."""
    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    rec = mll_step_record(args, w, dev, world, rank, steps=args.steps, warmup=args.warmup)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sym", action="store_true", help="use the SIMT forward kernel instead of the symmetric tensor-core kernel")
    ap.add_argument("--mode", default="cg", choices=["cg", "mll"], help="cg: one CG iteration per step (headline); mll: full MLL+gradient step")
    ap.add_argument("--cg-tol", type=float, default=0.002)
    ap.add_argument("--no-parity", action="store_true", help="skip the at-scale parity check against the FP64 C oracle")
    ap.add_argument("--parity-rows", type=int, default=256)
    ap.add_argument("--mll-workload", default="cfg2", choices=sorted(WORKLOADS) + ["none"],
                    help="workload of the `mll_step` sub-record (full MLL + gradient step through the model API); none = skip")
    ap.add_argument("--mll-steps", type=int, default=2)
    ap.add_argument("--mll-max-cg", type=int, default=0, help="cap on CG iterations of the MLL step (0 = the reference's 10 000)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 on their own (NCCL prints its
    # version there when NCCL_DEBUG is set in the environment) are sent to stderr while the benchmark runs
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved_stdout, "w")
    sys.stdout = real_stdout
    rc = 0
    try:
        if args.mode == "mll" and args.impl == "ours":
            rc = run_mll(args, w)
        elif args.impl == "reference":
            run_reference(args, w)
        else:
            rc = run_ours(args, w)
    finally:
        real_stdout.flush()
    return rc or 0


if __name__ == "__main__":
    sys.exit(main())

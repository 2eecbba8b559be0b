"""ctypes binding of librpgp.so -- the C ABI declared in include/rpgp.h.

The library is built in-tree (csrc/Makefile -> rpgp/librpgp.so).  There is NO fallback: if the shared object is
missing, or a call is made on a machine without a CUDA device, a RuntimeError is raised.  Tensors are owned by
PyTorch; only raw device pointers and the current CUDA stream cross the boundary.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RPGP_LIB") or os.path.join(_HERE, "librpgp.so")     # RPGP_LIB: an experimental build (tools/build_variant.sh)

LN2 = 0.6931471805599453


class Layout(Structure):
    """mirror of rpgp_layout (include/rpgp.h)"""
    _fields_ = [("J", c_int), ("K", c_int), ("CP", c_int), ("nchunks", c_int), ("KP", c_int), ("G", c_int), ("base", c_int)]

    def key(self):
        return (self.J, self.K, self.CP, self.nchunks, self.KP, self.G)

    def __repr__(self):
        return "Layout(J=%d, K=%d, CP=%d, nchunks=%d, KP=%d, G=%d)" % self.key()


# name -> (restype, argtypes); every symbol include/rpgp.h declares
SIGNATURES = {
    "rpgp_version": (c_int, []),
    "rpgp_last_error": (c_char_p, []),
    "rpgp_launch_count": (ctypes.c_ulonglong, []),
    "rpgp_plan_layout": (c_int, [c_int, c_int, POINTER(Layout)]),
    "rpgp_plan_layout_base": (c_int, [c_int, c_int, c_int, POINTER(Layout)]),
    "rpgp_padded_rhs": (c_int, [POINTER(Layout), c_int, c_int]),
    "rpgp_max_rhs": (c_int, [POINTER(Layout), c_int]),
    "rpgp_coord_scale": (c_double, []),
    "rpgp_pack_coords_f32": (c_int, [c_void_p, c_int64, c_int64, POINTER(Layout), c_float, c_void_p, c_void_p]),
    "rpgp_pack_log2c_f32": (c_int, [c_void_p, POINTER(Layout), c_void_p, c_void_p]),
    "rpgp_project_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p, POINTER(Layout),
                                 c_float, c_void_p, c_void_p]),
    "rpgp_project_tc_supported": (c_int, [c_int, POINTER(Layout)]),
    "rpgp_project2_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p, POINTER(Layout),
                                  c_float, c_void_p, c_void_p, c_int64, c_void_p]),
    "rpgp_project_bwd_workspace_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "rpgp_project_bwd_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "rpgp_mvm_workspace_bytes": (c_size_t, [c_int64, c_int64, POINTER(Layout), c_int]),
    "rpgp_mvm_fwd_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, POINTER(Layout), c_void_p,
                                 c_void_p, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "rpgp_mvm_sym_workspace_bytes": (c_size_t, [c_int64, POINTER(Layout)]),
    "rpgp_mvm_sym_supported": (c_int, [POINTER(Layout), c_int]),
    "rpgp_mvm_sym_distance_bound": (ctypes.c_float, []),
    "rpgp_mvm_sym_distance_plan": (c_int, [POINTER(Layout), POINTER(c_int)]),
    "rpgp_mvm_sym_f32": (c_int, [c_void_p, c_int64, POINTER(Layout), c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                 c_int, c_void_p, c_size_t, c_void_p]),
    "rpgp_quad_workspace_bytes": (c_size_t, [c_int64, c_int64, POINTER(Layout), c_int]),
    "rpgp_quad_bwd_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, POINTER(Layout), c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "rpgp_kernel_rows_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                     c_int64, c_void_p]),
    "rpgp_kernel_rows_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                     c_int64, c_void_p]),
    "rpgp_mvm_fwd_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                 c_int, c_void_p, c_void_p]),
    "rpgp_kernel_rows_base_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                          c_int64, c_void_p]),
    "rpgp_kernel_rows_base_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                          c_int64, c_void_p]),
    "rpgp_mvm_fwd_base_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_int, c_void_p, c_void_p]),
    "rpgp_quad_bwd_base_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "rpgp_quad_bwd_f64": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "rpgp_kmv_host_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int]),
    "rpgp_measure_peaks": (c_int, [POINTER(c_double), c_int, POINTER(c_char_p)]),
    "rpgp_plan_create": (c_int, [c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p, POINTER(c_void_p)]),
    "rpgp_plan_destroy": (c_int, [c_void_p]),
    "rpgp_plan_set_operator": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rpgp_plan_kmv_begin": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int]),
    "rpgp_plan_device_out": (c_void_p, [c_void_p]),
    "rpgp_plan_kmv_end": (c_int, [c_void_p, c_float, c_void_p, c_int64, c_int64]),
}

_lib = None


def load():
    """Load librpgp.so (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "librpgp.so not found at %s -- build it with `make -C %s` (or __graft_entry__.build()); "
            "there is no CPU fallback for the K.V path" % (LIB_PATH, os.path.join(os.path.dirname(_HERE), "csrc")))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        msg = load().rpgp_last_error()
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "rpgp: the K.V path runs only on CUDA tensors (sm_100a kernels in librpgp.so); got a %s tensor. "
                "There is no CPU fallback." % t.device)


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


_layout_cache = {}


BASE_RBF, BASE_MATERN15, BASE_INVERSE_MQ, BASE_COSINE = 0, 1, 2, 3      # rpgp_base_kernel (include/rpgp.h)


def plan_layout(J, K, base=BASE_RBF):
    key = (int(J), int(K), int(base))
    lay = _layout_cache.get(key)
    if lay is None:
        lay = Layout()
        _check(load().rpgp_plan_layout_base(key[0], key[1], key[2], ctypes.byref(lay)), "rpgp_plan_layout_base")
        _layout_cache[key] = lay
    return lay


def launch_count():
    return int(load().rpgp_launch_count())


def coord_scale():
    return float(load().rpgp_coord_scale())


def padded_rhs(lay, t, backward=False):
    return int(load().rpgp_padded_rhs(ctypes.byref(lay), int(t), int(bool(backward))))


def max_rhs(lay, backward=False):
    return int(load().rpgp_max_rhs(ctypes.byref(lay), int(bool(backward))))


# ---- workspace cache: one growing buffer per (device, stream), borrowed by the library during a launch --------------------
# (launches on different streams may overlap, so they must not share scratch memory)
_workspaces = {}


def _workspace(device, nbytes):
    if nbytes == 0:
        return None, 0
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf, buf.numel()


def free_workspaces():
    _workspaces.clear()


# ---- optional per-call device timing (bench.py's kernel split of an MLL step) ----------------------------------------------
_timing = None


class timing:
    """with _lib.timing() as tm: ... ; tm.totals() -> {entry point: (calls, ms)} from CUDA events recorded around every library
    call made on the current stream inside the block (synchronises once, when totals() is read)."""

    def __enter__(self):
        global _timing
        self.records, self._prev = [], _timing
        _timing = self
        return self

    def __exit__(self, *exc):
        global _timing
        _timing = self._prev

    def totals(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.records:
            calls, ms = out.get(name, (0, 0.0))
            out[name] = (calls + 1, ms + e0.elapsed_time(e1))
        return out


class _timed:
    __slots__ = ("name", "e1")

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _timing is not None:
            e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _timing.records.append((self.name, e0, self.e1))
            e0.record()
        return self

    def __exit__(self, *exc):
        if _timing is not None:
            self.e1.record()


# ---- thin wrappers ------------------------------------------------------------------------------------------------------
def pack_coords(Z, lay, scale=None):
    """natural (n x J*K) float32 CUDA tensor -> packed planes (nchunks, n, CP), scaled."""
    require_cuda(Z)
    assert Z.dtype == torch.float32 and Z.dim() == 2 and Z.stride(1) == 1
    n = Z.shape[0]
    out = torch.empty((lay.nchunks, n, lay.CP), dtype=torch.float32, device=Z.device)
    with torch.cuda.device(Z.device):
        _check(load().rpgp_pack_coords_f32(_ptr(Z), n, Z.stride(0), ctypes.byref(lay),
                                           coord_scale() if scale is None else float(scale), _ptr(out),
                                           _stream(Z.device)), "rpgp_pack_coords_f32")
    return out


def pack_log2c(c, lay):
    require_cuda(c)
    c = c.contiguous().float()
    out = torch.empty((lay.nchunks * lay.G,), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _check(load().rpgp_pack_log2c_f32(_ptr(c), ctypes.byref(lay), _ptr(out), _stream(c.device)), "rpgp_pack_log2c_f32")
    return out


def project(X, W, pre_inv, post_inv, lay, scale=None):
    """packed scaled projections of X (n x d) through W ((J*K) x d)."""
    require_cuda(X, W, pre_inv, post_inv)
    assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1
    W = W.contiguous().float()
    n, d = X.shape
    assert W.shape == (lay.J * lay.K, d)
    out = torch.empty((lay.nchunks, n, lay.CP), dtype=torch.float32, device=X.device)
    pre = None if pre_inv is None else pre_inv.contiguous().float()
    post = None if post_inv is None else post_inv.contiguous().float()
    with torch.cuda.device(X.device):
        _check(load().rpgp_project_f32(_ptr(X), n, d, X.stride(0), _ptr(W), _ptr(pre), _ptr(post), ctypes.byref(lay),
                                       coord_scale() if scale is None else float(scale), _ptr(out),
                                       _stream(X.device)), "rpgp_project_f32")
    return out


def project_tc_supported(d, lay):
    return bool(load().rpgp_project_tc_supported(int(d), ctypes.byref(lay)))


def project2(X, W, pre_inv, post_inv, lay, packed=True, natural=True, scale=None):
    """The projection with both outputs (rpgp_project2_f32): (packed planes or None, natural n x J*K float32 or None)."""
    require_cuda(X, W, pre_inv, post_inv)
    assert X.dtype == torch.float32 and X.dim() == 2 and X.stride(1) == 1
    W = W.contiguous().float()
    n, d = X.shape
    JK = lay.J * lay.K
    assert W.shape == (JK, d)
    zp = torch.empty((lay.nchunks, n, lay.CP), dtype=torch.float32, device=X.device) if packed else None
    zn = torch.empty((n, JK), dtype=torch.float32, device=X.device) if natural else None
    pre = None if pre_inv is None else pre_inv.reshape(-1).contiguous().float()
    post = None if post_inv is None else post_inv.reshape(-1).contiguous().float()
    with torch.cuda.device(X.device), _timed("project"):
        _check(load().rpgp_project2_f32(_ptr(X), n, d, X.stride(0), _ptr(W), _ptr(pre), _ptr(post), ctypes.byref(lay),
                                        coord_scale() if scale is None else float(scale), _ptr(zp), _ptr(zn), JK,
                                        _stream(X.device)), "rpgp_project2_f32")
    return zp, zn


def project_bwd(X, dZ):
    """dW'[q, k] = sum_i dZ[i, q] X[i, k]  (J*K x d): the vector-Jacobian product of the projection (rpgp_project_bwd_f32)."""
    require_cuda(X, dZ)
    assert X.dtype == torch.float32 and dZ.dtype == torch.float32 and X.dim() == 2 and dZ.dim() == 2
    assert X.stride(1) == 1 and X.shape[0] == dZ.shape[0]
    dZ = dZ if dZ.stride(1) == 1 else dZ.contiguous()
    n, d = X.shape
    JK = dZ.shape[1]
    lib = load()
    out = torch.empty((JK, d), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device), _timed("project_bwd"):
        ws, ws_bytes = _workspace(X.device, lib.rpgp_project_bwd_workspace_bytes(n, d, JK))
        _check(lib.rpgp_project_bwd_f32(_ptr(X), n, d, X.stride(0), _ptr(dZ), dZ.stride(0), JK, _ptr(out), _ptr(ws), ws_bytes,
                                        _stream(X.device)), "rpgp_project_bwd_f32")
    return out


def pad_rhs(V, TP):
    """(n x t) -> contiguous (n x TP), zero padded."""
    n, t = V.shape
    if t == TP and V.is_contiguous():
        return V
    out = torch.zeros((n, TP), dtype=V.dtype, device=V.device)
    out[:, :t] = V
    return out


def mvm_fwd(z1p, z2p, lay, nlc, V, row_range=None, events=None):
    """out = K(z1 rows, z2) @ V for packed planes.  row_range=(r0, r1) restricts to a row block of z1p.
    events=(start, stop): optional torch.cuda.Event pair recorded tightly around the library call (kernel timing)."""
    require_cuda(z1p, z2p, nlc, V)
    assert z1p.dtype == torch.float32 and z2p.dtype == torch.float32 and V.dtype == torch.float32
    assert z1p.is_contiguous() and z2p.is_contiguous() and nlc.is_contiguous()
    lib = load()
    m_full, n = z1p.shape[1], z2p.shape[1]
    r0, r1 = (0, m_full) if row_range is None else row_range
    m = r1 - r0
    t = V.shape[1]
    assert V.shape[0] == n
    out = torch.empty((m, t), dtype=torch.float32, device=V.device)
    tmax = max_rhs(lay, False)
    z1_ptr = c_void_p(z1p.data_ptr() + r0 * lay.CP * 4)
    with torch.cuda.device(V.device), _timed("mvm_fwd"):
        st = _stream(V.device)
        for t0 in range(0, t, tmax):
            tc = min(tmax, t - t0)
            TP = padded_rhs(lay, tc, False)
            Vp = pad_rhs(V[:, t0:t0 + tc], TP)
            nbytes = lib.rpgp_mvm_workspace_bytes(m, n, ctypes.byref(lay), tc)
            ws, ws_bytes = _workspace(V.device, nbytes)
            o = out[:, t0:t0 + tc]
            if events is not None and t0 == 0:
                events[0].record()
            _check(lib.rpgp_mvm_fwd_f32(z1_ptr, m, m_full * lay.CP, _ptr(z2p), n, n * lay.CP, ctypes.byref(lay),
                                        _ptr(nlc), _ptr(Vp), tc, _ptr(o), out.stride(0), _ptr(ws), ws_bytes, st),
                   "rpgp_mvm_fwd_f32")
        if events is not None:
            events[1].record()
    return out


def mvm_sym_supported(lay, t):
    return bool(load().rpgp_mvm_sym_supported(ctypes.byref(lay), int(t)))


def mvm_sym_distance_plan(lay):
    """chunking of the distance-on-tensor-core path for K > 1 (csrc/sym_tcd.cu), or None when the layout does not use it"""
    plan = (c_int * 5)()
    _check(load().rpgp_mvm_sym_distance_plan(ctypes.byref(lay), plan), "rpgp_mvm_sym_distance_plan")
    if not plan[0]:
        return None
    return dict(groups_per_chunk=plan[1], ksteps_per_group=plan[2], lines=plan[3], nchunks=plan[4],
                bound=float(load().rpgp_mvm_sym_distance_bound()))


def mvm_sym(zp, lay, nlc, V, block_range=None, events=None):
    """Symmetric K(Z,Z) @ V on the tensor cores (every kernel value evaluated once).  block_range=(b0, b1) restricts to the
    unique block pairs owned by 128-row blocks [b0, b1): the result then holds partial sums for ALL rows (all-reduce it)."""
    require_cuda(zp, nlc, V)
    assert zp.dtype == torch.float32 and zp.is_contiguous() and zp.shape[0] == lay.nchunks
    lib = load()
    n, t = zp.shape[1], V.shape[1]
    assert V.shape[0] == n
    nblocks = (n + 127) // 128
    b0, b1 = (0, nblocks) if block_range is None else block_range
    out = torch.empty((n, t), dtype=torch.float32, device=V.device)
    with torch.cuda.device(V.device), _timed("mvm_sym"):
        st = _stream(V.device)
        for t0 in range(0, t, 16):
            tc = min(16, t - t0)
            Vp = pad_rhs(V[:, t0:t0 + tc].float(), 16)
            nbytes = lib.rpgp_mvm_sym_workspace_bytes(n, ctypes.byref(lay))
            ws, ws_bytes = _workspace(V.device, nbytes)
            o = out[:, t0:t0 + tc]
            if events is not None and t0 == 0:
                events[0].record()
            _check(lib.rpgp_mvm_sym_f32(_ptr(zp), n, ctypes.byref(lay), _ptr(nlc), _ptr(Vp), tc, _ptr(o), out.stride(0),
                                        int(b0), int(b1), _ptr(ws), ws_bytes, st), "rpgp_mvm_sym_f32")
        if events is not None:
            events[1].record()
    return out


def quad_bwd(z1p, z2p, lay, nlc, L, R, symmetric, row_range=None, L_full=None, R_full=None):
    """Row-side gradient of sum_col L[:,col]^T K R[:,col].

    Non-symmetric: L is (m x t) for the rows of z1p, R is (n x t) for z2p.
    Symmetric (z1p is z2p): L, R are the full (n x t) vectors; row_range picks the row block.
    Returns (dz1p (nchunks, m, CP), g (nchunks*G,)).
    """
    require_cuda(z1p, z2p, nlc, L, R)
    lib = load()
    m_full, n = z1p.shape[1], z2p.shape[1]
    r0, r1 = (0, m_full) if row_range is None else row_range
    m = r1 - r0
    t = L.shape[1]
    tmax = max_rhs(lay, True)
    dz = None
    g = None
    z1_ptr = c_void_p(z1p.data_ptr() + r0 * lay.CP * 4)
    with torch.cuda.device(L.device), _timed("quad_bwd"):
        st = _stream(L.device)
        for t0 in range(0, t, tmax):
            tc = min(tmax, t - t0)
            TP = padded_rhs(lay, tc, True)
            if symmetric:
                Lc = pad_rhs(L[:, t0:t0 + tc], TP)
                Rc = pad_rhs(R[:, t0:t0 + tc], TP)
                Lrow = c_void_p(Lc.data_ptr() + r0 * TP * 4)
                Rrow = c_void_p(Rc.data_ptr() + r0 * TP * 4)
                Rcol, Lcol = _ptr(Rc), _ptr(Lc)
            else:
                Lr = pad_rhs(L[:, t0:t0 + tc], TP)
                Rc = pad_rhs(R[:, t0:t0 + tc], TP)
                assert Lr.shape[0] == m_full
                Lrow = c_void_p(Lr.data_ptr() + r0 * TP * 4)
                Rrow, Lcol = None, None
                Rcol = _ptr(Rc)
            nbytes = lib.rpgp_quad_workspace_bytes(m, n, ctypes.byref(lay), tc)
            ws, ws_bytes = _workspace(L.device, nbytes)
            dz_c = torch.empty((lay.nchunks, m, lay.CP), dtype=torch.float32, device=L.device)
            g_c = torch.empty((lay.nchunks * lay.G,), dtype=torch.float32, device=L.device)
            _check(lib.rpgp_quad_bwd_f32(z1_ptr, m, m_full * lay.CP, _ptr(z2p), n, n * lay.CP, ctypes.byref(lay),
                                         _ptr(nlc), Lrow, Rrow, Rcol, Lcol, tc, int(bool(symmetric)), _ptr(dz_c),
                                         _ptr(g_c), _ptr(ws), ws_bytes, st), "rpgp_quad_bwd_f32")
            dz = dz_c if dz is None else dz.add_(dz_c)
            g = g_c if g is None else g.add_(g_c)
    return dz, g


def kernel_rows(Zr, Z2, c, J, K, base=BASE_RBF):
    """dense K(Zr, Z2) (P x n) on natural coordinates, float32 or float64."""
    require_cuda(Zr, Z2, c)
    assert Zr.dtype == Z2.dtype and Zr.dtype in (torch.float32, torch.float64)
    Zr = Zr.contiguous()
    Z2 = Z2.contiguous()
    c = c.to(Zr.dtype).contiguous()
    P, n = Zr.shape[0], Z2.shape[0]
    out = torch.empty((P, n), dtype=Zr.dtype, device=Zr.device)
    fn = load().rpgp_kernel_rows_base_f32 if Zr.dtype == torch.float32 else load().rpgp_kernel_rows_base_f64
    with torch.cuda.device(Zr.device):
        _check(fn(_ptr(Zr), P, _ptr(Z2), n, J * K, J, K, int(base), _ptr(c), _ptr(out), n, _stream(Zr.device)), "rpgp_kernel_rows")
    return out


def mvm_fwd_f64(Z1, Z2, c, J, K, V, base=BASE_RBF):
    require_cuda(Z1, Z2, c, V)
    Z1, Z2, V = Z1.contiguous(), Z2.contiguous(), V.contiguous()
    c = c.to(torch.float64).contiguous()
    m, n, t = Z1.shape[0], Z2.shape[0], V.shape[1]
    out = torch.empty((m, t), dtype=torch.float64, device=V.device)
    with torch.cuda.device(V.device):
        _check(load().rpgp_mvm_fwd_base_f64(_ptr(Z1), m, _ptr(Z2), n, J * K, J, K, int(base), _ptr(c), _ptr(V), t, _ptr(out),
                                            _stream(V.device)), "rpgp_mvm_fwd_f64")
    return out


def quad_bwd_f64(Z1, Z2, c, J, K, L, R, base=BASE_RBF):
    """(dG/dZ1, g = dG/d ln c) for the FP64 path."""
    require_cuda(Z1, Z2, c, L, R)
    Z1, Z2, L, R = Z1.contiguous(), Z2.contiguous(), L.contiguous(), R.contiguous()
    c = c.to(torch.float64).contiguous()
    m, n, t = Z1.shape[0], Z2.shape[0], L.shape[1]
    dZ1 = torch.zeros_like(Z1)
    g = torch.zeros((J,), dtype=torch.float64, device=Z1.device)
    with torch.cuda.device(Z1.device):
        _check(load().rpgp_quad_bwd_base_f64(_ptr(Z1), m, _ptr(Z2), n, J * K, J, K, int(base), _ptr(c), _ptr(L), _ptr(R), t,
                                             _ptr(dZ1), _ptr(g), _stream(Z1.device)), "rpgp_quad_bwd_f64")
    return dZ1, g


def kmv_host(X1, X2, W, J, K, pre_inv, post_inv, c, V, diag_add=0.0, device=0):
    """Whole path on HOST numpy float32 arrays through rpgp_kmv_host_f32 (H2D + project + K.V + D2H)."""
    import numpy as np

    def arr(a):
        return None if a is None else np.ascontiguousarray(a, dtype=np.float32)

    X1, X2, W, pre_inv, post_inv, c, V = map(arr, (X1, X2, W, pre_inv, post_inv, c, V))
    m, d = X1.shape
    n = m if X2 is None else X2.shape[0]
    t = V.shape[1]
    out = np.empty((m, t), dtype=np.float32)

    def p(a):
        return None if a is None else a.ctypes.data_as(c_void_p)

    _check(load().rpgp_kmv_host_f32(p(X1), m, p(X2), n, d, p(W), J, K, p(pre_inv), p(post_inv), p(c), p(V), t,
                                    float(diag_add), p(out), int(device)), "rpgp_kmv_host_f32")
    return out


class _DevicePointer:
    """a borrowed device buffer for torch.as_tensor (CUDA array interface, float32, C-contiguous)"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3,
                                         "strides": None}


class HostPlan:
    """The host-buffer path as a persistent plan (rpgp_plan_*, include/rpgp.h): device buffers allocated once; per step
    `set_operator` (H2D of X, W, scales, c + projection), `kmv` (H2D of V, symmetric product of this rank's block pairs,
    all-reduce over the ranks of the default process group when there are several, + sigma^2 V, D2H of the rank's rows).
    numpy float32 arrays in, numpy out; pinned arrays (torch `.pin_memory().numpy()`) make the copies full-speed."""

    def __init__(self, n, d, J, K, tmax, device=0):
        self.n, self.d, self.J, self.K, self.tmax, self.device = int(n), int(d), int(J), int(K), int(tmax), int(device)
        self._h = c_void_p()
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.current_stream(self.device)
            _check(load().rpgp_plan_create(self.n, self.d, self.J, self.K, self.tmax, self.device,
                                           c_void_p(self._stream.cuda_stream), ctypes.byref(self._h)), "rpgp_plan_create")
        self._keep = None

    def close(self):
        if self._h:
            load().rpgp_plan_destroy(self._h)
            self._h = c_void_p()

    __del__ = close

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data_as(c_void_p)

    @staticmethod
    def _arr(a, shape=None):
        import numpy as np
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=np.float32)
        if shape is not None and tuple(a.shape) != tuple(shape):
            raise ValueError("expected shape %s, got %s" % (tuple(shape), tuple(a.shape)))
        return a

    def set_operator(self, X, W, pre_inv, post_inv, c):
        X, W, c = self._arr(X, (self.n, self.d)), self._arr(W, (self.J * self.K, self.d)), self._arr(c, (self.J,))
        pre_inv, post_inv = self._arr(pre_inv), self._arr(post_inv)
        self._keep = (X, W, pre_inv, post_inv, c)       # asynchronous copies: keep the host arrays alive
        _check(load().rpgp_plan_set_operator(self._h, self._p(X), self._p(W), self._p(pre_inv), self._p(post_inv), self._p(c)),
               "rpgp_plan_set_operator")

    def kmv(self, V, diag_add=0.0, block_range=None, row_range=None, all_reduce=None, out=None):
        """K(X,X) V + diag_add V.  block_range: this rank's 128-row blocks (default: all); all_reduce: callable applied to
        the device product (a torch view of the plan's buffer) between the kernel and the copy back; row_range: rows returned;
        out: optional (rows x t) float32 array to receive them (pinned: full-speed copy)."""
        import numpy as np
        V = self._arr(V)
        t = V.shape[1]
        nblocks = (self.n + 127) // 128
        b0, b1 = (0, nblocks) if block_range is None else block_range
        r0, r1 = (0, self.n) if row_range is None else row_range
        lib = load()
        _check(lib.rpgp_plan_kmv_begin(self._h, self._p(V), t, int(b0), int(b1)), "rpgp_plan_kmv_begin")
        if all_reduce is not None:
            with torch.cuda.device(self.device):
                dev_out = torch.as_tensor(_DevicePointer(lib.rpgp_plan_device_out(self._h), (self.n, t)),
                                          device=torch.device("cuda", self.device))
                all_reduce(dev_out)
        if out is None:
            out = np.empty((r1 - r0, t), dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == (r1 - r0, t) and out.flags["C_CONTIGUOUS"]
        _check(lib.rpgp_plan_kmv_end(self._h, float(diag_add), self._p(out), int(r0), int(r1)), "rpgp_plan_kmv_end")
        return out


def measure_peaks():
    """Run the issue-rate microbenchmarks; returns {name: {fp32_per_clk_sm, mufu_per_clk_sm, ms, mhz}}."""
    out = (c_double * 64)()
    names = (c_char_p * 16)()
    n = load().rpgp_measure_peaks(out, 16, names)
    if n < 0:
        raise RuntimeError("rpgp_measure_peaks failed: %d" % n)
    res = {}
    for i in range(n):
        res[names[i].decode()] = {"fp32_per_clk_sm": out[4 * i], "mufu_per_clk_sm": out[4 * i + 1],
                                  "ms": out[4 * i + 2], "mhz": out[4 * i + 3]}
    return res

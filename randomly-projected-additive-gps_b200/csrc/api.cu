// api.cu -- extern "C" entry points of librpgp.so (declared in include/rpgp.h): argument checks, layout planning,
// grid/split selection, workspace carving and the host-buffer variant of the whole path.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>

#include "../../include/rpgp.h"
#include "aux_kernels.cuh"
#include "dispatch.cuh"
#include "sym_tc.cuh"

namespace rpgp {
const char* last_error();

static_assert(sizeof(Layout) == sizeof(rpgp_layout), "Layout must mirror rpgp_layout");

// target number of CTAs per launch: enough waves on 148 SMs x 2 resident CTAs that the tail is a few percent
constexpr long long TARGET_CTAS = 148LL * 2 * 24;

struct SplitPlan {
    long long row_blocks, cols_per_split;
    int nsplits;
};

static SplitPlan plan_splits(long long m, long long n, int nchunks) {
    SplitPlan p;
    p.row_blocks = (m + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
    const long long tiles = (n + TN - 1) / TN;
    long long want = (TARGET_CTAS + p.row_blocks * nchunks - 1) / std::max<long long>(1, p.row_blocks * nchunks);
    // keep at least 4 tiles per split for big problems; allow single-tile splits when the problem is small
    const long long min_tiles = (p.row_blocks * nchunks * tiles >= 4 * TARGET_CTAS) ? 4 : 1;
    want = std::max<long long>(1, std::min<long long>(want, std::max<long long>(1, tiles / min_tiles)));
    want = std::min<long long>(want, 1024);
    long long tiles_per_split = (tiles + want - 1) / want;
    p.cols_per_split = tiles_per_split * TN;
    p.nsplits = (int)std::max<long long>(1, (n + p.cols_per_split - 1) / p.cols_per_split);
    return p;
}

static int check_layout(const rpgp_layout* lay) {
    RPGP_REQUIRE(lay != nullptr, "layout is NULL");
    RPGP_REQUIRE(lay->J >= 1 && lay->K >= 1, "layout: J=%d K=%d must be >= 1", lay->J, lay->K);
    RPGP_REQUIRE(lay->CP >= 4 && lay->CP <= 32 && lay->CP % 4 == 0, "layout: CP=%d must be a multiple of 4 in [4,32]", lay->CP);
    RPGP_REQUIRE(lay->nchunks >= 1 && lay->G >= 1 && lay->KP >= 1, "layout: nchunks/G/KP must be >= 1");
    RPGP_REQUIRE(lay->KP >= lay->K && lay->G * lay->KP <= lay->CP, "layout: group shape KP=%d G=%d exceeds CP=%d", lay->KP, lay->G, lay->CP);
    RPGP_REQUIRE((long long)lay->nchunks * lay->G >= lay->J, "layout: %d chunks x %d groups < J=%d", lay->nchunks, lay->G, lay->J);
    RPGP_REQUIRE(lay->base >= 0 && lay->base <= 3, "layout: base kernel %d (0 RBF, 1 Matern-1.5, 2 inverse multiquadric, 3 cosine)", lay->base);
    RPGP_REQUIRE((lay->K == 1) == (lay->KP == 1), "layout: KP must be 1 exactly when K is 1");
    RPGP_REQUIRE(lay->KP > 1 || lay->G == lay->CP, "layout: K=1 requires G == CP");
    return OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace rpgp

using namespace rpgp;

extern "C" {

int rpgp_version(void) { return 100; }
const char* rpgp_last_error(void) { return rpgp::last_error(); }
double rpgp_coord_scale(void) { return COORD_SCALE_D; }
unsigned long long rpgp_launch_count(void) { return rpgp::launch_count(); }

// TP actually compiled for (layout, t): forward K=1 {4,8,12,16,32}; backward K=1 {4,12,16}; K>1 {4,16}
int rpgp_padded_rhs(const rpgp_layout* lay, int t, int backward) {
    if (!lay || t <= 0) return 0;
    if (lay->KP > 1 || lay->base != 0) return t <= 4 ? 4 : (t <= 16 ? 16 : 0);
    if (backward) return t <= 4 ? 4 : (t <= 12 ? 12 : (t <= 16 ? 16 : 0));
    if (t <= 4) return 4;
    if (t <= 8) return 8;
    if (t <= 12) return 12;
    if (t <= 16) return 16;
    if (t <= 32) return 32;
    return 0;
}
int rpgp_max_rhs(const rpgp_layout* lay, int backward) { return (lay && lay->KP == 1 && lay->base == 0 && !backward) ? 32 : 16; }

int rpgp_plan_layout(int J, int K, rpgp_layout* out) { return rpgp_plan_layout_base(J, K, 0, out); }

int rpgp_plan_layout_base(int J, int K, int base, rpgp_layout* out) {
    RPGP_REQUIRE(out != nullptr, "plan_layout: out is NULL");
    RPGP_REQUIRE(J >= 1 && K >= 1, "plan_layout: J=%d K=%d must be >= 1", J, K);
    RPGP_REQUIRE(base >= 0 && base <= 3, "plan_layout: base kernel %d (0 RBF, 1 Matern-1.5, 2 inverse multiquadric, 3 cosine)", base);
    out->J = J;
    out->K = K;
    out->base = base;
    if (K == 1) {      // one coordinate per projection, every base kernel (round 2: the non-RBF kernels used to be padded to KP = 2)
        out->KP = 1;
        out->nchunks = (J + 31) / 32;
        const int per = (J + out->nchunks - 1) / out->nchunks;
        out->CP = ((per + 3) / 4) * 4;
        out->G = out->CP;
        return OK;
    }
    // (KP, G, CP) shapes compiled in dispatch.cuh (RPGP_KN_SHAPE_LIST)
    static const int shapes[][3] = {{2, 16, 32}, {2, 4, 8},  {4, 8, 32},  {4, 2, 8},   {6, 5, 32},  {6, 2, 12}, {8, 4, 32},
                                    {8, 1, 8},   {12, 2, 24}, {16, 2, 32}, {16, 1, 16}, {20, 1, 20}, {24, 1, 24}, {32, 1, 32}};
    long long best_cost = -1;
    for (const auto& s : shapes) {
        if (s[0] < K) continue;
        const int chunks = (J + s[1] - 1) / s[1];
        // cost ~ FP32 lane-ops + MUFU-equivalent per (i,i') pair: chunks * (G*(KP+1) + 16); every chunk is also a full pass of the
        // symmetric kernel's S.V / S^T.V products and of the right-hand sides (the +48)
        const long long cost = (long long)chunks * (s[1] * (s[0] + 8) + 16 + 48);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            out->KP = s[0];
            out->G = s[1];
            out->CP = s[2];
            out->nchunks = chunks;
        }
    }
    if (best_cost < 0) {
        set_error("plan_layout: K=%d exceeds the largest compiled group width (32)", K);
        return ERR_UNSUPPORTED;
    }
    return OK;
}

int rpgp_pack_coords_f32(const float* Z, int64_t n, int64_t ld, const rpgp_layout* lay, float scale, float* Zp, void* stream) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(n >= 0 && ld >= (int64_t)lay->J * lay->K, "pack_coords: n=%lld ld=%lld", (long long)n, (long long)ld);
    RPGP_REQUIRE(n == 0 || (Z && Zp), "pack_coords: NULL pointer");
    return launch_pack_coords(Z, n, ld, *reinterpret_cast<const Layout*>(lay), scale, Zp, (cudaStream_t)stream);
}

int rpgp_pack_log2c_f32(const float* c, const rpgp_layout* lay, float* neg_log2c, void* stream) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(c && neg_log2c, "pack_log2c: NULL pointer");
    return launch_pack_log2c(c, *reinterpret_cast<const Layout*>(lay), neg_log2c, (cudaStream_t)stream);
}

int rpgp_project_f32(const float* X, int64_t n, int d, int64_t ldx, const float* W, const float* pre_inv,
                     const float* post_inv, const rpgp_layout* lay, float scale, float* Zp, void* stream) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(n >= 0 && d >= 1 && ldx >= d, "project: n=%lld d=%d ldx=%lld", (long long)n, d, (long long)ldx);
    RPGP_REQUIRE(n == 0 || (X && W && Zp), "project: NULL pointer");
    return launch_project(X, n, d, ldx, W, pre_inv, post_inv, *reinterpret_cast<const Layout*>(lay), scale, Zp, nullptr, 0,
                          (cudaStream_t)stream);
}

int rpgp_project2_f32(const float* X, int64_t n, int d, int64_t ldx, const float* W, const float* pre_inv, const float* post_inv,
                      const rpgp_layout* lay, float scale, float* Zp, float* Zn, int64_t ldz, void* stream) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(n >= 0 && d >= 1 && ldx >= d, "project2: n=%lld d=%d ldx=%lld", (long long)n, d, (long long)ldx);
    RPGP_REQUIRE(n == 0 || (X && W && (Zp || Zn)), "project2: NULL pointer");
    RPGP_REQUIRE(Zn == nullptr || ldz >= (int64_t)lay->J * lay->K, "project2: ldz=%lld < J*K", (long long)ldz);
    return launch_project(X, n, d, ldx, W, pre_inv, post_inv, *reinterpret_cast<const Layout*>(lay), scale, Zp, Zn, ldz, (cudaStream_t)stream);
}

int rpgp_project_tc_supported(int d, const rpgp_layout* lay) {
    return lay && check_layout(lay) == OK && project_tc_supported(d, *reinterpret_cast<const Layout*>(lay));
}

size_t rpgp_project_bwd_workspace_bytes(int64_t n, int d, int JK) {
    if (n <= 0 || d <= 0 || JK <= 0) return 0;
    return project_bwd_workspace_bytes(n, d, JK);
}

int rpgp_project_bwd_f32(const float* X, int64_t n, int d, int64_t ldx, const float* dZ, int64_t ldz, int JK, float* dW, void* workspace,
                         size_t workspace_bytes, void* stream) {
    RPGP_REQUIRE(n >= 0 && d >= 1 && JK >= 1 && ldx >= d && ldz >= JK, "project_bwd: n=%lld d=%d JK=%d ldx=%lld ldz=%lld", (long long)n, d, JK,
                 (long long)ldx, (long long)ldz);
    RPGP_REQUIRE(dW != nullptr && (n == 0 || (X && dZ)), "project_bwd: NULL pointer");
    return launch_project_bwd(X, n, d, ldx, dZ, ldz, JK, dW, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t rpgp_mvm_workspace_bytes(int64_t m, int64_t n, const rpgp_layout* lay, int t) {
    if (!lay || m <= 0 || n <= 0 || t <= 0) return 0;
    const SplitPlan sp = plan_splits(m, n, lay->nchunks);
    const long long nparts = (long long)sp.nsplits * lay->nchunks;
    if (nparts == 1) return 0;
    const int TP = rpgp_padded_rhs(lay, std::min(t, rpgp_max_rhs(lay, 0)), 0);
    return (size_t)nparts * (size_t)m * TP * sizeof(float);
}

int rpgp_mvm_fwd_f32(const float* z1p, int64_t m, int64_t z1_stride, const float* z2p, int64_t n, int64_t z2_stride,
                     const rpgp_layout* lay, const float* neg_log2c, const float* Vp, int t, float* out, int ldo,
                     void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(m >= 0 && n >= 0 && t >= 1 && t <= rpgp_max_rhs(lay, 0),
                 "mvm_fwd: m=%lld n=%lld t=%d (1 <= t <= %d per call; chunk wider right-hand sides)", (long long)m,
                 (long long)n, t, rpgp_max_rhs(lay, 0));
    RPGP_REQUIRE(ldo >= t, "mvm_fwd: ldo=%d < t=%d", ldo, t);
    if (m == 0) return OK;
    RPGP_REQUIRE(out != nullptr, "mvm_fwd: out is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {  // empty sum
        RPGP_CUDA_OK(cudaMemset2DAsync(out, (size_t)ldo * sizeof(float), 0, (size_t)t * sizeof(float), (size_t)m, st));
        return OK;
    }
    RPGP_REQUIRE(z1p && z2p && neg_log2c && Vp, "mvm_fwd: NULL pointer");
    RPGP_REQUIRE(aligned16(z1p) && aligned16(z2p) && aligned16(Vp), "mvm_fwd: packed operands must be 16-byte aligned");
    RPGP_REQUIRE(z1_stride >= m * lay->CP && z2_stride >= n * lay->CP && z2_stride % 4 == 0,
                 "mvm_fwd: plane strides too small / unaligned");
    const int TP = rpgp_padded_rhs(lay, t, 0);
    const SplitPlan sp = plan_splits(m, n, lay->nchunks);
    const long long nparts = (long long)sp.nsplits * lay->nchunks;
    MvmArgs a;
    a.z1 = z1p; a.z2 = z2p; a.v = Vp; a.nlc = neg_log2c; a.out = out; a.partial = (float*)workspace;
    a.m = m; a.n = n; a.z1_chunk_stride = z1_stride; a.z2_chunk_stride = z2_stride;
    a.cols_per_split = sp.cols_per_split; a.ldo = ldo; a.t = t; a.nsplits = sp.nsplits; a.nchunks = lay->nchunks;
    a.direct = (nparts == 1);
    if (!a.direct) {
        const size_t need = (size_t)nparts * (size_t)m * TP * sizeof(float);
        if (workspace == nullptr || workspace_bytes < need) {
            set_error("mvm_fwd: workspace %zu bytes < required %zu", workspace_bytes, need);
            return ERR_WORKSPACE;
        }
    }
    dim3 grid((unsigned)sp.row_blocks, (unsigned)sp.nsplits, (unsigned)lay->nchunks);
    // RPGP_POLY_PAIRS overrides the number of projection pairs evaluated on the FMA pipe (tools/poly_sweep.py); -1 = default
    static const int poly_env = [] { const char* e = getenv("RPGP_POLY_PAIRS"); return e ? atoi(e) : -1; }();
    const int poly_pairs = (lay->KP == 1 && lay->base == 0) ? (poly_env >= 0 ? poly_env : default_poly_pairs(lay->CP, TP)) : 0;
    int rc;
    if (lay->KP == 1 && lay->base != 0) rc = launch_fwd_k1_base(lay->CP, TP, lay->base, a, grid, st);
    else if (lay->KP == 1 && poly_pairs > 0) rc = launch_fwd_k1_poly(lay->CP, TP, poly_pairs, a, grid, st);
    else rc = (lay->KP == 1) ? launch_fwd_k1(lay->CP, TP, a, grid, st) : launch_fwd_kn(lay->KP, lay->G, lay->CP, TP, lay->base, a, grid, st);
    if (rc) return rc;
    if (!a.direct) return launch_reduce_partials(a.partial, (int)nparts, m, TP, t, out, ldo, st);
    return OK;
}

size_t rpgp_mvm_sym_workspace_bytes(int64_t n, const rpgp_layout* lay) {
    if (!lay || n <= 0) return 0;
    return sym_workspace_bytes(n, *reinterpret_cast<const Layout*>(lay));
}

float rpgp_mvm_sym_distance_bound(void) { return tcd_gate_bound(); }

int rpgp_mvm_sym_distance_plan(const rpgp_layout* lay, int plan[5]) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(plan != nullptr, "mvm_sym_distance_plan: NULL pointer");
    const TcdPlan p = plan_tcd(*reinterpret_cast<const Layout*>(lay));
    plan[0] = p.supported; plan[1] = p.GT; plan[2] = p.KS; plan[3] = p.NL; plan[4] = p.nchunks;
    return OK;
}

int rpgp_mvm_sym_supported(const rpgp_layout* lay, int t) {
    // every base kernel for K = 1; K > 1: RBF, Matern-1.5, inverse MQ (the cosine kernel with K > 1 stays on the rectangular kernel; the
    // distance-on-tensor-core variant is RBF only)
    return lay && check_layout(lay) == OK && t >= 1 && t <= 16 && !(lay->base == 3 && lay->KP > 1);
}

int rpgp_mvm_sym_f32(const float* zp, int64_t n, const rpgp_layout* lay, const float* neg_log2c, const float* Vp16, int t,
                     float* out, int ldo, int row_block_begin, int row_block_end, void* workspace, size_t workspace_bytes,
                     void* stream) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(rpgp_mvm_sym_supported(lay, t), "mvm_sym: needs 1 <= t <= 16 right-hand sides per call (got t=%d) and a layout the symmetric kernel takes (base %d, KP %d: see rpgp_mvm_sym_supported)", t, lay ? lay->base : -1, lay ? lay->KP : -1);
    RPGP_REQUIRE(n >= 1 && ldo >= t, "mvm_sym: n=%lld ldo=%d", (long long)n, ldo);
    RPGP_REQUIRE(zp && neg_log2c && Vp16 && out, "mvm_sym: NULL pointer");
    RPGP_REQUIRE(aligned16(zp) && aligned16(Vp16), "mvm_sym: operands must be 16-byte aligned");
    const int nblocks = (int)((n + 127) / 128);
    RPGP_REQUIRE(0 <= row_block_begin && row_block_begin <= row_block_end && row_block_end <= nblocks,
                 "mvm_sym: row block range [%d, %d) outside [0, %d]", row_block_begin, row_block_end, nblocks);
    return launch_sym_tc5(zp, n, *reinterpret_cast<const Layout*>(lay), neg_log2c, Vp16, t, out, ldo, row_block_begin, row_block_end, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

size_t rpgp_quad_workspace_bytes(int64_t m, int64_t n, const rpgp_layout* lay, int t) {
    if (!lay || m <= 0 || n <= 0 || t <= 0) return 0;
    const SplitPlan sp = plan_splits(m, n, lay->nchunks);
    const size_t dz = (size_t)sp.nsplits * lay->nchunks * (size_t)m * lay->CP * sizeof(float);
    const size_t g = (size_t)sp.row_blocks * sp.nsplits * lay->nchunks * lay->G * sizeof(float);
    return dz + ((g + 255) / 256) * 256;
}

int rpgp_quad_bwd_f32(const float* z1p, int64_t m, int64_t z1_stride, const float* z2p, int64_t n, int64_t z2_stride,
                      const rpgp_layout* lay, const float* neg_log2c, const float* Lrow, const float* Rrow,
                      const float* Rcol, const float* Lcol, int t, int symmetric, float* dz1p, float* g,
                      void* workspace, size_t workspace_bytes, void* stream) {
    if (int rc = check_layout(lay)) return rc;
    RPGP_REQUIRE(m >= 0 && n >= 0 && t >= 1 && t <= 16, "quad_bwd: m=%lld n=%lld t=%d (1 <= t <= 16 per call)", (long long)m, (long long)n, t);
    RPGP_REQUIRE(dz1p && g, "quad_bwd: output pointer is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const int width = lay->nchunks * lay->G;
    if (m == 0 || n == 0) {
        if (m > 0) RPGP_CUDA_OK(cudaMemsetAsync(dz1p, 0, (size_t)lay->nchunks * m * lay->CP * sizeof(float), st));
        RPGP_CUDA_OK(cudaMemsetAsync(g, 0, (size_t)width * sizeof(float), st));
        return OK;
    }
    RPGP_REQUIRE(z1p && z2p && neg_log2c && Lrow && Rcol, "quad_bwd: NULL pointer");
    RPGP_REQUIRE(!symmetric || (Rrow && Lcol), "quad_bwd: symmetric mode needs Rrow and Lcol");
    RPGP_REQUIRE(aligned16(z1p) && aligned16(z2p) && aligned16(Rcol) && (!symmetric || aligned16(Lcol)),
                 "quad_bwd: packed operands must be 16-byte aligned");
    RPGP_REQUIRE(z1_stride >= m * lay->CP && z2_stride >= n * lay->CP && z2_stride % 4 == 0, "quad_bwd: plane strides too small / unaligned");
    const int TPk = rpgp_padded_rhs(lay, t, 1);
    const SplitPlan sp = plan_splits(m, n, lay->nchunks);
    const size_t need = rpgp_quad_workspace_bytes(m, n, lay, t);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("quad_bwd: workspace %zu bytes < required %zu", workspace_bytes, need);
        return ERR_WORKSPACE;
    }
    const size_t dz_bytes = (size_t)sp.nsplits * lay->nchunks * (size_t)m * lay->CP * sizeof(float);
    GradArgs a;
    a.z1 = z1p; a.z2 = z2p; a.a_row = Lrow; a.b_row = Rrow; a.r_col = Rcol; a.l_col = Lcol; a.nlc = neg_log2c;
    a.dz_partial = (float*)workspace;
    a.g_partial = (float*)((char*)workspace + dz_bytes);
    a.m = m; a.n = n; a.z1_chunk_stride = z1_stride; a.z2_chunk_stride = z2_stride;
    a.cols_per_split = sp.cols_per_split; a.nsplits = sp.nsplits; a.nchunks = lay->nchunks; a.symmetric = symmetric ? 1 : 0;
    dim3 grid((unsigned)sp.row_blocks, (unsigned)sp.nsplits, (unsigned)lay->nchunks);
    int rc = (lay->KP == 1 && lay->base != 0) ? launch_grad_k1_base(lay->CP, TPk, lay->base, a, grid, st)
             : (lay->KP == 1)                 ? launch_grad_k1(lay->CP, TPk, a, grid, st)
                                              : launch_grad_kn(lay->KP, lay->G, lay->CP, TPk, lay->base, a, grid, st);
    if (rc) return rc;
    // d k / d z1 = k * ln2 * (-2 d)  in scaled coordinates
    const float dz_scale = (float)(-2.0 * LN2_D);
    // dz partial planes are [split][chunk][m][CP]; dz1p is written contiguous: [chunk][m][CP]
    rc = launch_reduce_dz(a.dz_partial, sp.nsplits, (long long)lay->nchunks * m * lay->CP, dz_scale, dz1p, st);
    if (rc) return rc;
    return launch_reduce_g(a.g_partial, sp.row_blocks * sp.nsplits, width, symmetric ? 0.5f : 1.0f, g, st);
}

int rpgp_kernel_rows_base_f32(const float* Zr, int64_t P, const float* Z2, int64_t n, int64_t ld, int J, int K, int base,
                              const float* c, float* out, int64_t ldo, void* stream) {
    RPGP_REQUIRE(P >= 0 && n >= 0 && J >= 1 && K >= 1 && ld >= (int64_t)J * K && ldo >= n, "kernel_rows: bad shape");
    RPGP_REQUIRE(base >= 0 && base <= 3, "kernel_rows: base kernel %d", base);
    RPGP_REQUIRE((P == 0 || n == 0) || (Zr && Z2 && c && out), "kernel_rows: NULL pointer");
    return launch_rows_f32(Zr, P, Z2, n, ld, J, K, base, c, out, ldo, (cudaStream_t)stream);
}
int rpgp_kernel_rows_base_f64(const double* Zr, int64_t P, const double* Z2, int64_t n, int64_t ld, int J, int K, int base,
                              const double* c, double* out, int64_t ldo, void* stream) {
    RPGP_REQUIRE(P >= 0 && n >= 0 && J >= 1 && K >= 1 && ld >= (int64_t)J * K && ldo >= n, "kernel_rows: bad shape");
    RPGP_REQUIRE(base >= 0 && base <= 3, "kernel_rows: base kernel %d", base);
    RPGP_REQUIRE((P == 0 || n == 0) || (Zr && Z2 && c && out), "kernel_rows: NULL pointer");
    return launch_rows_f64(Zr, P, Z2, n, ld, J, K, base, c, out, ldo, (cudaStream_t)stream);
}
int rpgp_kernel_rows_f32(const float* Zr, int64_t P, const float* Z2, int64_t n, int64_t ld, int J, int K,
                         const float* c, float* out, int64_t ldo, void* stream) {
    return rpgp_kernel_rows_base_f32(Zr, P, Z2, n, ld, J, K, 0, c, out, ldo, stream);
}
int rpgp_kernel_rows_f64(const double* Zr, int64_t P, const double* Z2, int64_t n, int64_t ld, int J, int K,
                         const double* c, double* out, int64_t ldo, void* stream) {
    return rpgp_kernel_rows_base_f64(Zr, P, Z2, n, ld, J, K, 0, c, out, ldo, stream);
}

int rpgp_mvm_fwd_base_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K, int base,
                          const double* c, const double* V, int t, double* out, void* stream) {
    RPGP_REQUIRE(m >= 0 && n >= 0 && J >= 1 && K >= 1 && t >= 1 && ld >= (int64_t)J * K, "mvm_fwd_f64: bad shape");
    RPGP_REQUIRE(base >= 0 && base <= 3, "mvm_fwd_f64: base kernel %d", base);
    RPGP_REQUIRE(m == 0 || (Z1 && c && out && (n == 0 || (Z2 && V))), "mvm_fwd_f64: NULL pointer");
    return launch_mvm_f64(Z1, m, Z2, n, ld, J, K, base, c, V, t, out, (cudaStream_t)stream);
}
int rpgp_quad_bwd_base_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K, int base,
                           const double* c, const double* L, const double* R, int t, double* dZ1, double* g, void* stream) {
    RPGP_REQUIRE(m >= 0 && n >= 0 && J >= 1 && K >= 1 && t >= 1 && ld >= (int64_t)J * K, "quad_bwd_f64: bad shape");
    RPGP_REQUIRE(base >= 0 && base <= 3, "quad_bwd_f64: base kernel %d", base);
    RPGP_REQUIRE(m == 0 || n == 0 || (Z1 && Z2 && c && L && R && dZ1 && g), "quad_bwd_f64: NULL pointer");
    if (n == 0) return OK;
    return launch_quad_f64(Z1, m, Z2, n, ld, J, K, base, c, L, R, t, dZ1, g, (cudaStream_t)stream);
}
int rpgp_mvm_fwd_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K,
                     const double* c, const double* V, int t, double* out, void* stream) {
    return rpgp_mvm_fwd_base_f64(Z1, m, Z2, n, ld, J, K, 0, c, V, t, out, stream);
}
int rpgp_quad_bwd_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K,
                      const double* c, const double* L, const double* R, int t, double* dZ1, double* g, void* stream) {
    return rpgp_quad_bwd_base_f64(Z1, m, Z2, n, ld, J, K, 0, c, L, R, t, dZ1, g, stream);
}

// ---- whole path on host buffers ---------------------------------------------------------------------------------------
int rpgp_kmv_host_f32(const float* X1, int64_t m, const float* X2, int64_t n, int d, const float* W, int J, int K,
                      const float* pre_inv, const float* post_inv, const float* c, const float* V, int t,
                      float diag_add, float* out, int device) {
    RPGP_REQUIRE(X1 && W && c && V && out, "kmv_host: NULL pointer");
    RPGP_REQUIRE(m >= 1 && d >= 1 && J >= 1 && K >= 1 && t >= 1, "kmv_host: bad shape");
    const bool square = (X2 == nullptr);
    if (square) n = m;
    RPGP_REQUIRE(n >= 1, "kmv_host: n must be >= 1");
    RPGP_CUDA_OK(cudaSetDevice(device));
    rpgp_layout lay;
    if (int rc = rpgp_plan_layout(J, K, &lay)) return rc;
    const int JK = J * K;
    // square products of a few thousand rows or more go to the symmetric tensor-core kernel (16 right-hand sides per pass)
    const bool use_sym = square && m >= 1024 && rpgp_mvm_sym_supported(&lay, std::min(t, 16));
    const int tmax = use_sym ? 16 : rpgp_max_rhs(&lay, 0);
    const float scale = (float)COORD_SCALE_D;

    struct Buffers {
        std::vector<void*> ptrs;
        cudaStream_t st = nullptr;
        ~Buffers() {
            for (void* p : ptrs) cudaFree(p);
            if (st) cudaStreamDestroy(st);
        }
        int alloc(void** p, size_t bytes) {
            cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
            ptrs.push_back(*p);
            return OK;
        }
    } buf;
    RPGP_CUDA_OK(cudaStreamCreateWithFlags(&buf.st, cudaStreamNonBlocking));
    cudaStream_t st = buf.st;
    // RPGP_HOST_TIMING=1: wall-clock phases of this call on stderr (adds a synchronisation after every phase)
    static const bool timing = [] { const char* e = getenv("RPGP_HOST_TIMING"); return e && atoi(e) != 0; }();
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    double t_prev = now();
    auto phase = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(st);
        const double t_now = now();
        fprintf(stderr, "[rpgp_kmv_host] %-28s %9.3f ms\n", what, 1e3 * (t_now - t_prev));
        t_prev = t_now;
    };
    float *dX1, *dX2 = nullptr, *dW, *dpre = nullptr, *dpost = nullptr, *dc, *dnlc, *dZ1, *dZ2, *dV, *dVp, *dout;
    void* dws;
    int rc;
    if ((rc = buf.alloc((void**)&dX1, (size_t)m * d * 4))) return rc;
    if (!square && (rc = buf.alloc((void**)&dX2, (size_t)n * d * 4))) return rc;
    if ((rc = buf.alloc((void**)&dW, (size_t)JK * d * 4))) return rc;
    if (pre_inv && (rc = buf.alloc((void**)&dpre, (size_t)d * 4))) return rc;
    if (post_inv && (rc = buf.alloc((void**)&dpost, (size_t)JK * 4))) return rc;
    if ((rc = buf.alloc((void**)&dc, (size_t)J * 4))) return rc;
    if ((rc = buf.alloc((void**)&dnlc, (size_t)lay.nchunks * lay.G * 4))) return rc;
    if ((rc = buf.alloc((void**)&dZ1, (size_t)lay.nchunks * m * lay.CP * 4))) return rc;
    if (!square && (rc = buf.alloc((void**)&dZ2, (size_t)lay.nchunks * n * lay.CP * 4))) return rc;
    if ((rc = buf.alloc((void**)&dV, (size_t)n * t * 4))) return rc;
    const int tc0 = std::min(t, tmax);
    const int TP0 = use_sym ? 16 : rpgp_padded_rhs(&lay, tc0, 0);
    if ((rc = buf.alloc((void**)&dVp, (size_t)n * TP0 * 4))) return rc;
    if ((rc = buf.alloc((void**)&dout, (size_t)m * t * 4))) return rc;
    const size_t ws_bytes = use_sym ? rpgp_mvm_sym_workspace_bytes(n, &lay) : rpgp_mvm_workspace_bytes(m, n, &lay, tc0);
    if ((rc = buf.alloc(&dws, ws_bytes))) return rc;
    phase("allocations");

    RPGP_CUDA_OK(cudaMemcpyAsync(dX1, X1, (size_t)m * d * 4, cudaMemcpyHostToDevice, st));
    if (!square) RPGP_CUDA_OK(cudaMemcpyAsync(dX2, X2, (size_t)n * d * 4, cudaMemcpyHostToDevice, st));
    RPGP_CUDA_OK(cudaMemcpyAsync(dW, W, (size_t)JK * d * 4, cudaMemcpyHostToDevice, st));
    if (pre_inv) RPGP_CUDA_OK(cudaMemcpyAsync(dpre, pre_inv, (size_t)d * 4, cudaMemcpyHostToDevice, st));
    if (post_inv) RPGP_CUDA_OK(cudaMemcpyAsync(dpost, post_inv, (size_t)JK * 4, cudaMemcpyHostToDevice, st));
    RPGP_CUDA_OK(cudaMemcpyAsync(dc, c, (size_t)J * 4, cudaMemcpyHostToDevice, st));
    RPGP_CUDA_OK(cudaMemcpyAsync(dV, V, (size_t)n * t * 4, cudaMemcpyHostToDevice, st));

    phase("host -> device copies");
    if ((rc = rpgp_pack_log2c_f32(dc, &lay, dnlc, st))) return rc;
    if ((rc = rpgp_project_f32(dX1, m, d, d, dW, dpre, dpost, &lay, scale, dZ1, st))) return rc;
    if (!square && (rc = rpgp_project_f32(dX2, n, d, d, dW, dpre, dpost, &lay, scale, dZ2, st))) return rc;
    const float* z2 = square ? dZ1 : dZ2;
    phase("projection");
    for (int t0 = 0; t0 < t; t0 += tmax) {
        const int tc = std::min(tmax, t - t0);
        const int TP = use_sym ? 16 : rpgp_padded_rhs(&lay, tc, 0);
        // pad this block of right-hand sides to [n][TP]
        RPGP_CUDA_OK(cudaMemsetAsync(dVp, 0, (size_t)n * TP * 4, st));
        RPGP_CUDA_OK(cudaMemcpy2DAsync(dVp, (size_t)TP * 4, dV + t0, (size_t)t * 4, (size_t)tc * 4, (size_t)n,
                                       cudaMemcpyDeviceToDevice, st));
        if (use_sym)
            rc = rpgp_mvm_sym_f32(dZ1, n, &lay, dnlc, dVp, tc, dout + t0, t, 0, (int)((n + 127) / 128), dws, ws_bytes, st);
        else
            rc = rpgp_mvm_fwd_f32(dZ1, m, m * lay.CP, z2, n, n * lay.CP, &lay, dnlc, dVp, tc, dout + t0, t, dws, ws_bytes, st);
        if (rc) return rc;
    }
    if (square && diag_add != 0.f && (rc = launch_axpy_rows(diag_add, dV, t, m, t, dout, t, st))) return rc;
    phase("K.V");
    RPGP_CUDA_OK(cudaMemcpyAsync(out, dout, (size_t)m * t * 4, cudaMemcpyDeviceToHost, st));
    RPGP_CUDA_OK(cudaStreamSynchronize(st));
    phase("device -> host copy");
    return OK;
}


// ---- the host-buffer path as a persistent plan ------------------------------------------------------------------------------
struct rpgp_plan {
    long long n;
    int d, J, K, tmax, device;
    rpgp_layout lay;
    bool use_sym, own_stream, has_operator;
    cudaStream_t st;
    float *dX, *dW, *dpre, *dpost, *dc, *dnlc, *dZ, *dV, *dVp, *dout;
    void* dws;
    size_t ws_bytes;
    int t_last;
};

static void plan_free(rpgp_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (void* q : {(void*)p->dX, (void*)p->dW, (void*)p->dpre, (void*)p->dpost, (void*)p->dc, (void*)p->dnlc, (void*)p->dZ, (void*)p->dV,
                    (void*)p->dVp, (void*)p->dout, p->dws})
        if (q) cudaFree(q);
    if (p->own_stream && p->st) cudaStreamDestroy(p->st);
    delete p;
}

int rpgp_plan_create(int64_t n, int d, int J, int K, int tmax, int device, void* stream, rpgp_plan** out) {
    RPGP_REQUIRE(out != nullptr, "plan_create: out is NULL");
    *out = nullptr;
    RPGP_REQUIRE(n >= 1 && d >= 1 && J >= 1 && K >= 1 && tmax >= 1, "plan_create: bad shape n=%lld d=%d J=%d K=%d tmax=%d", (long long)n, d, J, K, tmax);
    RPGP_CUDA_OK(cudaSetDevice(device));
    rpgp_plan* p = new rpgp_plan();
    std::memset(p, 0, sizeof(*p));
    p->n = n; p->d = d; p->J = J; p->K = K; p->tmax = tmax; p->device = device;
    if (int rc = rpgp_plan_layout(J, K, &p->lay)) { delete p; return rc; }
    p->use_sym = n >= 1024 && rpgp_mvm_sym_supported(&p->lay, std::min(tmax, 16));
    p->own_stream = (stream == nullptr);
    p->st = (cudaStream_t)stream;
    auto fail = [&](int rc) { plan_free(p); return rc; };
    if (p->own_stream) {
        cudaError_t e = cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking);
        if (e != cudaSuccess) return fail(cuda_fail(e, "cudaStreamCreateWithFlags"));
    }
    const int JK = J * K;
    const int tper = p->use_sym ? 16 : rpgp_max_rhs(&p->lay, 0);
    const int TPmax = p->use_sym ? 16 : rpgp_padded_rhs(&p->lay, std::min(tmax, tper), 0);
    p->ws_bytes = p->use_sym ? rpgp_mvm_sym_workspace_bytes(n, &p->lay) : rpgp_mvm_workspace_bytes(n, n, &p->lay, std::min(tmax, tper));
    struct { void** ptr; size_t bytes; } allocs[] = {
        {(void**)&p->dX, (size_t)n * d * 4}, {(void**)&p->dW, (size_t)JK * d * 4}, {(void**)&p->dpre, (size_t)d * 4},
        {(void**)&p->dpost, (size_t)JK * 4}, {(void**)&p->dc, (size_t)J * 4}, {(void**)&p->dnlc, (size_t)p->lay.nchunks * p->lay.G * 4},
        {(void**)&p->dZ, (size_t)p->lay.nchunks * n * p->lay.CP * 4}, {(void**)&p->dV, (size_t)n * tmax * 4},
        {(void**)&p->dVp, (size_t)n * TPmax * 4}, {(void**)&p->dout, (size_t)n * tmax * 4}, {&p->dws, p->ws_bytes}};
    for (auto& a : allocs) {
        cudaError_t e = cudaMalloc(a.ptr, a.bytes ? a.bytes : 16);
        if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc (rpgp_plan_create)"));
    }
    *out = p;
    return OK;
}

int rpgp_plan_destroy(rpgp_plan* p) {
    plan_free(p);
    return OK;
}

int rpgp_plan_set_operator(rpgp_plan* p, const float* X, const float* W, const float* pre_inv, const float* post_inv, const float* c) {
    RPGP_REQUIRE(p && X && W && c, "plan_set_operator: NULL pointer");
    RPGP_CUDA_OK(cudaSetDevice(p->device));
    cudaStream_t st = p->st;
    const int JK = p->J * p->K;
    RPGP_CUDA_OK(cudaMemcpyAsync(p->dX, X, (size_t)p->n * p->d * 4, cudaMemcpyHostToDevice, st));
    RPGP_CUDA_OK(cudaMemcpyAsync(p->dW, W, (size_t)JK * p->d * 4, cudaMemcpyHostToDevice, st));
    if (pre_inv) RPGP_CUDA_OK(cudaMemcpyAsync(p->dpre, pre_inv, (size_t)p->d * 4, cudaMemcpyHostToDevice, st));
    if (post_inv) RPGP_CUDA_OK(cudaMemcpyAsync(p->dpost, post_inv, (size_t)JK * 4, cudaMemcpyHostToDevice, st));
    RPGP_CUDA_OK(cudaMemcpyAsync(p->dc, c, (size_t)p->J * 4, cudaMemcpyHostToDevice, st));
    if (int rc = rpgp_pack_log2c_f32(p->dc, &p->lay, p->dnlc, st)) return rc;
    if (int rc = rpgp_project_f32(p->dX, p->n, p->d, p->d, p->dW, pre_inv ? p->dpre : nullptr, post_inv ? p->dpost : nullptr, &p->lay,
                                  (float)COORD_SCALE_D, p->dZ, st))
        return rc;
    p->has_operator = true;
    return OK;
}

int rpgp_plan_kmv_begin(rpgp_plan* p, const float* V, int t, int rb_begin, int rb_end) {
    RPGP_REQUIRE(p && V, "plan_kmv_begin: NULL pointer");
    RPGP_REQUIRE(p->has_operator, "plan_kmv_begin: rpgp_plan_set_operator has not been called");
    RPGP_REQUIRE(t >= 1 && t <= p->tmax, "plan_kmv_begin: t=%d outside [1, tmax=%d]", t, p->tmax);
    const long long n = p->n;
    const int nblocks = (int)((n + 127) / 128);
    RPGP_REQUIRE(0 <= rb_begin && rb_begin <= rb_end && rb_end <= nblocks, "plan_kmv_begin: row block range [%d, %d) outside [0, %d]", rb_begin,
                 rb_end, nblocks);
    RPGP_CUDA_OK(cudaSetDevice(p->device));
    cudaStream_t st = p->st;
    RPGP_CUDA_OK(cudaMemcpyAsync(p->dV, V, (size_t)n * t * 4, cudaMemcpyHostToDevice, st));
    const int tper = p->use_sym ? 16 : rpgp_max_rhs(&p->lay, 0);
    const long long r0 = std::min<long long>(n, 128LL * rb_begin), r1 = std::min<long long>(n, 128LL * rb_end);
    if (!p->use_sym && (r0 > 0 || r1 < n)) RPGP_CUDA_OK(cudaMemsetAsync(p->dout, 0, (size_t)n * t * 4, st));
    for (int t0 = 0; t0 < t; t0 += tper) {
        const int tc = std::min(tper, t - t0);
        const int TP = p->use_sym ? 16 : rpgp_padded_rhs(&p->lay, tc, 0);
        RPGP_CUDA_OK(cudaMemsetAsync(p->dVp, 0, (size_t)n * TP * 4, st));
        RPGP_CUDA_OK(cudaMemcpy2DAsync(p->dVp, (size_t)TP * 4, p->dV + t0, (size_t)t * 4, (size_t)tc * 4, (size_t)n, cudaMemcpyDeviceToDevice, st));
        int rc;
        if (p->use_sym)
            rc = rpgp_mvm_sym_f32(p->dZ, n, &p->lay, p->dnlc, p->dVp, tc, p->dout + t0, t, rb_begin, rb_end, p->dws, p->ws_bytes, st);
        else if (r1 > r0) {
            const size_t need = rpgp_mvm_workspace_bytes(r1 - r0, n, &p->lay, tc);      // the split plan depends on the block height
            if (need > p->ws_bytes) {
                RPGP_CUDA_OK(cudaStreamSynchronize(st));
                if (p->dws) cudaFree(p->dws);
                p->dws = nullptr;
                p->ws_bytes = 0;
                RPGP_CUDA_OK(cudaMalloc(&p->dws, need));
                p->ws_bytes = need;
            }
            rc = rpgp_mvm_fwd_f32(p->dZ + r0 * p->lay.CP, r1 - r0, n * p->lay.CP, p->dZ, n, n * p->lay.CP, &p->lay, p->dnlc, p->dVp, tc,
                                  p->dout + r0 * t + t0, t, p->dws, p->ws_bytes, st);
        } else
            rc = OK;
        if (rc) return rc;
    }
    p->t_last = t;
    return OK;
}

void* rpgp_plan_device_out(rpgp_plan* p) { return p ? (void*)p->dout : nullptr; }

int rpgp_plan_kmv_end(rpgp_plan* p, float diag_add, float* out, int64_t row_begin, int64_t row_end) {
    RPGP_REQUIRE(p && out, "plan_kmv_end: NULL pointer");
    RPGP_REQUIRE(p->t_last >= 1, "plan_kmv_end: no product in flight (call rpgp_plan_kmv_begin first)");
    RPGP_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= p->n, "plan_kmv_end: rows [%lld, %lld) outside [0, %lld]",
                 (long long)row_begin, (long long)row_end, p->n);
    RPGP_CUDA_OK(cudaSetDevice(p->device));
    cudaStream_t st = p->st;
    const int t = p->t_last;
    if (diag_add != 0.f)
        if (int rc = launch_axpy_rows(diag_add, p->dV, t, p->n, t, p->dout, t, st)) return rc;
    if (row_end > row_begin)
        RPGP_CUDA_OK(cudaMemcpyAsync(out, p->dout + row_begin * t, (size_t)(row_end - row_begin) * t * 4, cudaMemcpyDeviceToHost, st));
    RPGP_CUDA_OK(cudaStreamSynchronize(st));
    p->t_last = 0;
    return OK;
}

}  // extern "C"

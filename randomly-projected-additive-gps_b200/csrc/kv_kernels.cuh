// kv_kernels.cuh -- the fused matrix-free kernel-matrix kernels for sm_100a.
//
//   forward   out[i,:]  = sum_{i'} K[i,i'] V[i',:]                                  
//   row-grad  dZ[i,q]   = d/dZ1[i,q] sum_{i,i'} S[i,i'] K[i,i'],  g_j = sum S k_j     
//   with      K[i,i']   = sum_j 2^( log2c_j - sum_{m<K} (z1[i,jK+m] - z2[i',jK+m])^2 )   (coordinates pre-scaled
//             by sqrt(log2(e)/2), see rpgp_common.cuh) -- formed tile by tile, never materialised.
//
// Replaces, behind the drop-in boundary, what the reference reaches through GPyTorch/KeOps
// (`SumLazyTensor._matmul` -> J x `KeOpsLazyTensor._matmul`; in-repo witness gp_models/kernels/imq_kernel.py:32-58)
// and the hand-written dense gradient of gp_models/kernels/memory_efficient_gam_kernel.py:33-59.
//
// Mapping onto the SM (B200: 148 SMs, MUFU 16/clk/SM, FP32 128 lanes/clk/SM -- measured by peaks.cu):
//   * one thread owns one row i: its CP coordinates, the -log2c constants and the t accumulators stay in registers
//     for the whole pass; a CTA is 256 rows.
//   * column tiles (TN columns of Z2 | V, contiguous in HBM because the operands are stored packed) are staged into
//     shared memory by the TMA engine as 1-D bulk copies (cp.async.bulk + mbarrier complete_tx), NSTAGE deep.
//     Every thread of a warp reads the same column -> LDS.128 broadcasts, no bank conflicts.
//   * K=1: pairs of projections are processed as packed f32x2: FADD2 (d), FFMA2 (d*d - log2c), 2x MUFU.EX2, FADD2.
//     K>1: pairs of coordinates inside a projection group are packed the same way (FADD2 + FFMA2), one MUFU per group.
//   * the V update is t/2 FFMA2 per pair; accumulation is two-level (per tile, then a compensated add into the
//     running total) so the FP32 error does not grow with n.
//   * grid = (row blocks, column splits, coordinate chunks); partial results go to a workspace and are summed in a
//     fixed order by reduce_partials_kernel -> bit-reproducible run to run.
#pragma once
#include "rpgp_common.cuh"

namespace rpgp {

constexpr int ROWS_PER_CTA = 256;
constexpr int TN = 64;      // columns per shared-memory tile
constexpr int NSTAGE = 4;   // TMA pipeline depth

struct MvmArgs {
    const float* z1;      // [nchunks][m][CP]   row-side packed coordinates (this rank's row block)
    const float* z2;      // [nchunks][n][CP]   column-side packed coordinates
    const float* v;       // [n][TP]            right-hand sides, zero padded to TP
    const float* nlc;     // [nchunks][GP]      -log2(c_j), +inf for padding groups
    float* out;           // final output [m][ldo] (used when no partials are needed)
    float* partial;       // [nparts][m][TP]
    long long m, n;
    long long z1_chunk_stride, z2_chunk_stride;  // elements between coordinate chunks
    long long cols_per_split;
    int ldo, t;
    int nsplits, nchunks;
    int direct;           // 1: single part -> write `out` directly
};

struct GradArgs {
    const float* z1;      // rows   [nchunks][m][CP]
    const float* z2;      // cols   [nchunks][n][CP]
    const float* a_row;   // [m][TP]   left vectors of the row block            (L)
    const float* b_row;   // [m][TP]   right vectors of the row block, symmetric mode only (R)
    const float* r_col;   // [n][TP]   right vectors on the column side          (R)
    const float* l_col;   // [n][TP]   left vectors on the column side, symmetric mode only (L)
    const float* nlc;     // [nchunks][GP]
    float* dz_partial;    // [nsplits][nchunks][m][CP]   (un-scaled: sum S k d ; the reducer applies -2 ln2)
    float* g_partial;     // [gridDim.x * nsplits][nchunks][GP]  per-CTA sums of S k_j
    long long m, n;
    long long z1_chunk_stride, z2_chunk_stride;
    long long cols_per_split;
    int nsplits, nchunks, symmetric;
};

#ifdef __CUDACC__

// ----------------------------------------------------------------------------------------------------------------
// per-pair kernel value:  s = sum_j 2^(lc_j - |dz_j|^2)
// ----------------------------------------------------------------------------------------------------------------
template <int CP, int KP, int G>
struct RowCoords {
    f32x2 z[CP / 2];                      // packed row coordinates
    f32x2 c2[(KP == 1) ? CP / 2 : 1];     // K=1: packed -log2c per projection pair (RBF) or packed weights c (other base kernels)
    float cg[(KP == 1) ? 1 : G];          // K>1: -log2c per group
    float cw[(KP == 1) ? 1 : G];          // K>1, non-RBF base kernels: c per group (the weight multiplies outside the base function)
};

// ---- base kernels (training_routines.py:57-83, imq_kernel.py:8-9,47) ---------------------------------------------
// u = scaled squared distance of a group (natural d^2 = 2 ln2 * u).  value: k = c f(d^2); slope: kz = -(dk/du) / ln2, the factor the
// row-gradient kernel multiplies the coordinate differences with (its reducer applies -2 ln2, which is exact for the RBF k = 2^-u).
constexpr int BASE_RBF = 0, BASE_MATERN15 = 1, BASE_IMQ = 2, BASE_COS = 3;
constexpr float MATERN_C = 4.1588830833596715f;      // 3 * 2 ln2: (sqrt3 r)^2 = MATERN_C * u
constexpr float TWO_LN2_F = 1.3862943611198906f;

// one MUFU each (sqrtf / rsqrtf expand to range checks and Newton steps; the .approx forms are accurate to ~2^-23)
__device__ __forceinline__ float sqrt_approx_ftz(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// cosine base kernel (gpytorch.kernels.CosineKernel as training_routines.py:76-81 selects it): k = c cos(d), d the natural distance of
// the group -- the kernel classes fold pi / period_length into the coordinates.  cos.approx / sin.approx are accurate to 2^-20.9 only
// near the origin, so the argument is first reduced to [-pi, pi] with a two-term Cody-Waite split of 2 pi (exact for |d| < 2^18).
__device__ __forceinline__ float reduce_2pi(float d) {
    // n = round(d / 2 pi) by the 1.5 * 2^23 magic constant: rintf is an FRND on the quarter-rate conversion pipe, which this kernel's
    // one MUFU per projection already fills (first version: 49 ms at the cfg2 shape against 26 ms for the other base kernels)
    const float n = __fadd_rn(fmaf(d, 0.15915494309189535f, 12582912.f), -12582912.f);
    float r = fmaf(n, -6.2831854820251465f, d);      // 2 pi rounded to FP32 ...
    return fmaf(n, 1.7484555314695172e-07f, r);     // ... and the remainder (2 pi = 6.2831854820251465 - 1.7484555e-7)
}
__device__ __forceinline__ float cos_approx_ftz(float x) {
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx_ftz(float x) {
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// sin(d) / d for the gradient (1 - d^2 / 6 near the origin, where the quotient loses its digits)
__device__ __forceinline__ float sinc_of(float d) {
    const float q = sin_approx_ftz(reduce_2pi(d)) * __frcp_rn(fmaxf(d, 1e-20f));
    return d < 0.02f ? fmaf(d * d, -0.16666667f, 1.f) : q;
}
constexpr float INV_SCALE_F = 1.1774100225154747f;   // natural coordinate = packed coordinate * sqrt(2 ln 2)

template <int BASE>
__device__ __forceinline__ float base_value(float u, float nl, float cw) {
    if constexpr (BASE == BASE_RBF) {
        return ex2_ftz(-(u + nl));
    } else if constexpr (BASE == BASE_MATERN15) {
        const float q = sqrt_approx_ftz(fmaxf(MATERN_C * u, 0.f));
        const float e = ex2_ftz(-q * LOG2E_F);
        return cw * fmaf(q, e, e);
    } else if constexpr (BASE == BASE_IMQ) {
        return cw * rsqrt_approx_ftz(fmaf(TWO_LN2_F, u, 1.f));
    } else {
        return cw * cos_approx_ftz(reduce_2pi(sqrt_approx_ftz(fmaxf(TWO_LN2_F * u, 0.f))));
    }
}

template <int BASE>
__device__ __forceinline__ void base_value_slope(float u, float nl, float cw, float& k, float& kz) {
    if constexpr (BASE == BASE_RBF) {
        k = ex2_ftz(-(u + nl));
        kz = k;
    } else if constexpr (BASE == BASE_MATERN15) {
        const float q = sqrt_approx_ftz(fmaxf(MATERN_C * u, 0.f));
        const float e = cw * ex2_ftz(-q * LOG2E_F);
        k = fmaf(q, e, e);
        kz = 3.f * e;                      // -dk/du = c (MATERN_C / 2) e^-q ; / ln2 = 3 c e^-q
    } else if constexpr (BASE == BASE_IMQ) {
        const float rs = rsqrt_approx_ftz(fmaf(TWO_LN2_F, u, 1.f));
        k = cw * rs;
        kz = k * rs * rs;                  // -dk/du = c ln2 rs^3
    } else {
        const float d = sqrt_approx_ftz(fmaxf(TWO_LN2_F * u, 0.f));
        k = cw * cos_approx_ftz(reduce_2pi(d));
        kz = cw * sinc_of(d);              // -dk/du = c sin(d) (2 ln2) / (2 d); / ln2 = c sin(d) / d
    }
}

// K = 1 (one coordinate per projection), non-RBF base kernels, on a packed pair of scaled differences d (natural d = ds / sqrt(log2(e)/2)):
// no square root is needed -- the distance is |d| -- so every base kernel costs ONE MUFU per projection, like the RBF.
//   Matern-1.5: q = sqrt3 |d_nat| = |ds| sqrt(MATERN_C);  a = q log2(e);  k = c 2^-a (1 + a ln2);  kz = -(dk/du)/ln2 = 3 c 2^-a
//   inverse MQ: k = c rsqrt(1 + 2 ln2 ds^2);  kz = k rs^2
//   cosine:     k = c cos(|d_nat|);  kz = c sin(|d_nat|) / |d_nat|   (two MUFU in the gradient: sin and the reciprocal)
constexpr float MATERN_A = 2.942137020149432f;  // sqrt(MATERN_C) * log2(e)
constexpr float LN2_F = 0.6931471805599453f;
template <int BASE>
__device__ __forceinline__ f32x2 base_value_k1(f32x2 d, f32x2 cw) {
    float dl, dh;
    unpack2(d, dl, dh);
    if constexpr (BASE == BASE_MATERN15) {
        const float al = fabsf(dl) * MATERN_A, ah = fabsf(dh) * MATERN_A;
        const f32x2 e = pack2(ex2_ftz(-al), ex2_ftz(-ah));
        const f32x2 u = fma2(pack2(al, ah), pack2(LN2_F, LN2_F), pack2(1.f, 1.f));
        return mul2(mul2(cw, e), u);
    } else if constexpr (BASE == BASE_IMQ) {
        const f32x2 t = fma2(d, mul2(d, pack2(TWO_LN2_F, TWO_LN2_F)), pack2(1.f, 1.f));
        float tl, th;
        unpack2(t, tl, th);
        return mul2(cw, pack2(rsqrt_approx_ftz(tl), rsqrt_approx_ftz(th)));
    } else {        // cosine: even in d, no absolute value needed
        return mul2(cw, pack2(cos_approx_ftz(reduce_2pi(dl * INV_SCALE_F)), cos_approx_ftz(reduce_2pi(dh * INV_SCALE_F))));
    }
}
template <int BASE>
__device__ __forceinline__ void base_value_slope_k1(f32x2 d, f32x2 cw, f32x2& k, f32x2& kz) {
    float dl, dh;
    unpack2(d, dl, dh);
    if constexpr (BASE == BASE_MATERN15) {
        const float al = fabsf(dl) * MATERN_A, ah = fabsf(dh) * MATERN_A;
        const f32x2 e = mul2(cw, pack2(ex2_ftz(-al), ex2_ftz(-ah)));
        k = mul2(e, fma2(pack2(al, ah), pack2(LN2_F, LN2_F), pack2(1.f, 1.f)));
        kz = mul2(e, pack2(3.f, 3.f));
    } else if constexpr (BASE == BASE_IMQ) {
        const f32x2 t = fma2(d, mul2(d, pack2(TWO_LN2_F, TWO_LN2_F)), pack2(1.f, 1.f));
        float tl, th;
        unpack2(t, tl, th);
        const f32x2 rs = pack2(rsqrt_approx_ftz(tl), rsqrt_approx_ftz(th));
        k = mul2(cw, rs);
        kz = mul2(k, mul2(rs, rs));
    } else {
        const float al = fabsf(dl) * INV_SCALE_F, ah = fabsf(dh) * INV_SCALE_F;
        k = mul2(cw, pack2(cos_approx_ftz(reduce_2pi(al)), cos_approx_ftz(reduce_2pi(ah))));
        kz = mul2(cw, pack2(sinc_of(al), sinc_of(ah)));
    }
}

// 2^(-u) for a packed pair on the FMA + ALU pipes instead of the XU pipe (takes load off MUFU, the binding unit):
// Cody-Waite split with the 1.5*2^23 magic constant (round-to-nearest integer n, |f| <= 1/2), degree-5 minimax polynomial
// for 2^f (max relative error 7.6e-8, i.e. the accuracy of ex2.approx), exponent inserted by an integer shift-add.
// u is clamped to <= 126 so that the biased exponent cannot wrap (results below 2^-126 flush towards 0 like .ftz).
__device__ __forceinline__ f32x2 exp2_neg_poly2(f32x2 u) {
    const f32x2 MAGIC = pack2(12582912.f, 12582912.f);
    // coefficients of q(g) = 2^(-g), g = -f  (odd powers negated)
    const f32x2 C5 = pack2(-0.0013280353741720319f, -0.0013280353741720319f);
    const f32x2 C4 = pack2(0.009675574488937855f, 0.009675574488937855f);
    const f32x2 C3 = pack2(-0.05550701171159744f, -0.05550701171159744f);
    const f32x2 C2 = pack2(0.24022118747234344f, 0.24022118747234344f);
    const f32x2 C1 = pack2(-0.6931470036506653f, -0.6931470036506653f);
    const f32x2 C0 = pack2(1.0000001192092896f, 1.0000001192092896f);
    float ul, uh;
    unpack2(u, ul, uh);
    u = pack2(fminf(ul, 126.f), fminf(uh, 126.f));
    const f32x2 rr = sub2(MAGIC, u);          // magic + x, x = -u : low mantissa bits hold n = round(x)
    const f32x2 nn = sub2(rr, MAGIC);         // n as float
    const f32x2 g = add2(u, nn);              // g = -(x - n) = -f
    f32x2 acc = fma2(C5, g, C4);
    acc = fma2(acc, g, C3);
    acc = fma2(acc, g, C2);
    acc = fma2(acc, g, C1);
    acc = fma2(acc, g, C0);
    float pl, ph, rl, rh;
    unpack2(acc, pl, ph);
    unpack2(rr, rl, rh);
    const float el = __int_as_float(__float_as_int(pl) + (__float_as_int(rl) << 23));
    const float eh = __int_as_float(__float_as_int(ph) + (__float_as_int(rh) << 23));
    return pack2(el, eh);
}

// NP2 = number of packed projection pairs per (i,i') whose exponential is evaluated by exp2_neg_poly2 (K = 1 only)
template <int CP, int KP, int G, int NP2 = 0, int BASE = 0>
__device__ __forceinline__ float pair_kernel_value(const RowCoords<CP, KP, G>& r, const float* __restrict__ zcol) {
    static_assert(BASE == 0 || NP2 == 0, "the polynomial exponential belongs to the RBF kernel");
    // zcol: CP floats of one column in shared memory (16 B aligned)
    f32x2 zj[CP / 2];
#pragma unroll
    for (int q = 0; q < CP / 4; ++q) {
        const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(zcol + 4 * q);
        zj[2 * q] = p.x;
        zj[2 * q + 1] = p.y;
    }
    if constexpr (KP == 1 && BASE != 0) {
        f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) {
            const f32x2 e = base_value_k1<BASE>(sub2(r.z[q], zj[q]), r.c2[q]);
            if (q & 1) s1 = add2(s1, e); else s0 = add2(s0, e);
        }
        float lo, hi;
        unpack2(add2(s0, s1), lo, hi);
        return lo + hi;
    } else if constexpr (KP == 1) {
        f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) {
            const f32x2 d = sub2(r.z[q], zj[q]);
            const f32x2 u = fma2(d, d, r.c2[q]);         // d^2 - log2c
            f32x2 e;
            if (q >= CP / 2 - NP2) {
                e = exp2_neg_poly2(u);
            } else {
                float ul, uh;
                unpack2(u, ul, uh);
                e = pack2(ex2_ftz(-ul), ex2_ftz(-uh));
            }
            if (q == 0) s0 = e;                          // two partial sums; no add for their first terms
            else if (q == 1) s1 = e;
            else if (q & 1) s1 = add2(s1, e);
            else s0 = add2(s0, e);
        }
        float lo, hi;
        if constexpr (CP / 2 >= 2) unpack2(add2(s0, s1), lo, hi); else unpack2(s0, lo, hi);
        return lo + hi;
    } else {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            f32x2 sq = pack2(BASE == 0 ? r.cg[g] : 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < KP / 2; ++p) {
                const f32x2 d = sub2(r.z[g * (KP / 2) + p], zj[g * (KP / 2) + p]);
                sq = fma2(d, d, sq);
            }
            float lo, hi;
            unpack2(sq, lo, hi);
            if constexpr (BASE == 0) s += ex2_ftz(-(lo + hi));
            else s += base_value<BASE>(lo + hi, 0.f, r.cw[g]);
        }
        return s;
    }
}

template <int CP, int KP, int G, int BASE = 0>
__device__ __forceinline__ void load_row_coords(RowCoords<CP, KP, G>& r, const float* __restrict__ zrow, bool valid,
                                                const float* __restrict__ nlc) {
#pragma unroll
    for (int q = 0; q < CP / 2; ++q) {
        float2 p = valid ? __ldg(reinterpret_cast<const float2*>(zrow) + q) : make_float2(0.f, 0.f);
        r.z[q] = pack2(p.x, p.y);
    }
    if constexpr (KP == 1 && BASE != 0) {      // weights c = 2^-(-log2 c); 0 for padding projections (+inf)
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) r.c2[q] = pack2(ex2_ftz(-__ldg(nlc + 2 * q)), ex2_ftz(-__ldg(nlc + 2 * q + 1)));
    } else if constexpr (KP == 1) {
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) r.c2[q] = pack2(__ldg(nlc + 2 * q), __ldg(nlc + 2 * q + 1));
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            r.cg[g] = __ldg(nlc + g);
            r.cw[g] = ex2_ftz(-r.cg[g]);        // c = 2^-(-log2 c); 0 for padding groups (+inf)
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// forward:  grid (row blocks, column splits, coordinate chunks), 256 threads, dynamic smem = fwd_smem_bytes<CP,TP>()
// ----------------------------------------------------------------------------------------------------------------
template <int CP, int TP>
constexpr size_t fwd_smem_bytes() { return 128 + (size_t)NSTAGE * TN * (CP + TP) * sizeof(float); }

template <int CP, int TP, int KP, int G, int NP2 = 0, int BASE = 0>
__global__ void __launch_bounds__(ROWS_PER_CTA, 2) mvm_fwd_kernel(const MvmArgs a) {
    static_assert(CP % 4 == 0 && TP % 4 == 0, "packed layouts (16 B rows for the bulk copies)");
    static_assert(KP == 1 || (KP % 2 == 0 && G * KP <= CP), "group layout");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    float* ztile = reinterpret_cast<float*>(smem_raw + 128);
    float* vtile = ztile + (size_t)NSTAGE * TN * CP;

    const int tid = threadIdx.x;
    const int chunk = blockIdx.z;
    const long long row = (long long)blockIdx.x * ROWS_PER_CTA + tid;
    const bool valid = row < a.m;
    const long long col0 = (long long)blockIdx.y * a.cols_per_split;
    const long long col1 = min(a.n, col0 + a.cols_per_split);
    const int ntiles = (int)((col1 - col0 + TN - 1) / TN);
    const float* z2 = a.z2 + (long long)chunk * a.z2_chunk_stride;
    constexpr int GP = (KP == 1) ? CP : G;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int tile) {
        const int s = tile % NSTAGE;
        const long long c0 = col0 + (long long)tile * TN;
        const uint32_t cols = (uint32_t)min((long long)TN, col1 - c0);
        mbar_expect_tx(&full[s], cols * (CP + TP) * (uint32_t)sizeof(float));
        bulk_g2s(ztile + (size_t)s * TN * CP, z2 + c0 * CP, cols * CP * (uint32_t)sizeof(float), &full[s]);
        bulk_g2s(vtile + (size_t)s * TN * TP, a.v + c0 * TP, cols * TP * (uint32_t)sizeof(float), &full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < NSTAGE && s < ntiles; ++s) issue(s);
    }

    RowCoords<CP, KP, G> r;
    load_row_coords<CP, KP, G, BASE>(r, a.z1 + (long long)chunk * a.z1_chunk_stride + row * CP, valid,
                                     a.nlc + (long long)chunk * GP);

    f32x2 acc[TP / 2], comp[TP / 2];
#pragma unroll
    for (int q = 0; q < TP / 2; ++q) { acc[q] = 0ull; comp[q] = 0ull; }

    for (int tile = 0; tile < ntiles; ++tile) {
        const int s = tile % NSTAGE;
        mbar_wait(&full[s], (uint32_t)((tile / NSTAGE) & 1));
        const int cols = (int)min((long long)TN, col1 - (col0 + (long long)tile * TN));
        const float* zt = ztile + (size_t)s * TN * CP;
        const float* vt = vtile + (size_t)s * TN * TP;
        f32x2 lo[TP / 2];
#pragma unroll
        for (int q = 0; q < TP / 2; ++q) lo[q] = 0ull;
#pragma unroll 2
        for (int c = 0; c < cols; ++c) {
            const float sv = pair_kernel_value<CP, KP, G, NP2, BASE>(r, zt + c * CP);
            const f32x2 ss = pack2(sv, sv);
#pragma unroll
            for (int q = 0; q < TP / 4; ++q) {
                const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(vt + c * TP + 4 * q);
                lo[2 * q] = fma2(ss, p.x, lo[2 * q]);
                lo[2 * q + 1] = fma2(ss, p.y, lo[2 * q + 1]);
            }
        }
        // compensated (Kahan) fold of the tile sum into the running total
#pragma unroll
        for (int q = 0; q < TP / 2; ++q) {
            const f32x2 y = sub2(lo[q], comp[q]);
            const f32x2 tsum = add2(acc[q], y);
            comp[q] = sub2(sub2(tsum, acc[q]), y);
            acc[q] = tsum;
        }
        __syncthreads();  // everyone is done with stage s
        if (tid == 0 && tile + NSTAGE < ntiles) issue(tile + NSTAGE);
    }

    if (!valid) return;
    float res[TP];
#pragma unroll
    for (int q = 0; q < TP / 2; ++q) unpack2(acc[q], res[2 * q], res[2 * q + 1]);
    if (a.direct) {
        float* o = a.out + row * a.ldo;
#pragma unroll
        for (int c = 0; c < TP; ++c)
            if (c < a.t) o[c] = res[c];
    } else {
        const long long part = (long long)chunk * a.nsplits + blockIdx.y;
        float4* o = reinterpret_cast<float4*>(a.partial + (part * a.m + row) * TP);
#pragma unroll
        for (int q = 0; q < TP / 4; ++q) o[q] = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// row-side gradient of the quadratic form.  S[i,i'] = A_i . R_i'  (+ B_i . L_i' in symmetric mode)
//   gz[q]  = sum_{i'} S k_j (z1[i,q] - z2[i',q])          (caller scales by -2 ln2 to get d/dz1)
//   gc[j]  = sum_{i'} S k_j                                (CTA-reduced, one partial row per CTA)
// ----------------------------------------------------------------------------------------------------------------
template <int CP, int TP>
constexpr size_t grad_smem_bytes(bool symmetric) {
    return 128 + (size_t)NSTAGE * TN * (CP + (symmetric ? 2 : 1) * TP) * sizeof(float) + 8 * 32 * sizeof(float);
}

template <int CP, int TP, int KP, int G, bool SYM, int BASE = 0>
__global__ void __launch_bounds__(ROWS_PER_CTA, 1) quad_rowgrad_kernel(const GradArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    float* ztile = reinterpret_cast<float*>(smem_raw + 128);
    float* rtile = ztile + (size_t)NSTAGE * TN * CP;
    float* ltile = rtile + (size_t)NSTAGE * TN * TP;  // SYM only
    float* red = rtile + (size_t)NSTAGE * TN * TP * (SYM ? 2 : 1);  // [8 warps][32]
    constexpr int GP = (KP == 1) ? CP : G;

    const int tid = threadIdx.x;
    const int chunk = blockIdx.z;
    const long long row = (long long)blockIdx.x * ROWS_PER_CTA + tid;
    const bool valid = row < a.m;
    const long long col0 = (long long)blockIdx.y * a.cols_per_split;
    const long long col1 = min(a.n, col0 + a.cols_per_split);
    const int ntiles = (int)((col1 - col0 + TN - 1) / TN);
    const float* z2 = a.z2 + (long long)chunk * a.z2_chunk_stride;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int tile) {
        const int s = tile % NSTAGE;
        const long long c0 = col0 + (long long)tile * TN;
        const uint32_t cols = (uint32_t)min((long long)TN, col1 - c0);
        mbar_expect_tx(&full[s], cols * (CP + (SYM ? 2 : 1) * TP) * (uint32_t)sizeof(float));
        bulk_g2s(ztile + (size_t)s * TN * CP, z2 + c0 * CP, cols * CP * (uint32_t)sizeof(float), &full[s]);
        bulk_g2s(rtile + (size_t)s * TN * TP, a.r_col + c0 * TP, cols * TP * (uint32_t)sizeof(float), &full[s]);
        if (SYM) bulk_g2s(ltile + (size_t)s * TN * TP, a.l_col + c0 * TP, cols * TP * (uint32_t)sizeof(float), &full[s]);
    };
    if (tid == 0) {
        for (int s = 0; s < NSTAGE && s < ntiles; ++s) issue(s);
    }

    RowCoords<CP, KP, G> r;
    load_row_coords<CP, KP, G, BASE>(r, a.z1 + (long long)chunk * a.z1_chunk_stride + row * CP, valid,
                                     a.nlc + (long long)chunk * GP);
    f32x2 arow[TP / 2], brow[SYM ? TP / 2 : 1];
#pragma unroll
    for (int q = 0; q < TP / 2; ++q) {
        const float2 p = valid ? __ldg(reinterpret_cast<const float2*>(a.a_row + row * TP) + q) : make_float2(0.f, 0.f);
        arow[q] = pack2(p.x, p.y);
        if (SYM) {
            const float2 pb = valid ? __ldg(reinterpret_cast<const float2*>(a.b_row + row * TP) + q) : make_float2(0.f, 0.f);
            brow[q] = pack2(pb.x, pb.y);
        }
    }
    f32x2 gz[CP / 2];
    f32x2 gc2[(KP == 1) ? CP / 2 : 1];
    float gcg[(KP == 1) ? 1 : G];
#pragma unroll
    for (int q = 0; q < CP / 2; ++q) gz[q] = 0ull;
    if constexpr (KP == 1) {
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) gc2[q] = 0ull;
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) gcg[g] = 0.f;
    }

    for (int tile = 0; tile < ntiles; ++tile) {
        const int s = tile % NSTAGE;
        mbar_wait(&full[s], (uint32_t)((tile / NSTAGE) & 1));
        const int cols = (int)min((long long)TN, col1 - (col0 + (long long)tile * TN));
        const float* zt = ztile + (size_t)s * TN * CP;
        const float* rt = rtile + (size_t)s * TN * TP;
        const float* lt = ltile + (size_t)s * TN * TP;
#pragma unroll 1
        for (int c = 0; c < cols; ++c) {
            // S = A_i . R_c (+ B_i . L_c)
            f32x2 s2 = 0ull;
#pragma unroll
            for (int q = 0; q < TP / 4; ++q) {
                const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(rt + c * TP + 4 * q);
                s2 = fma2(arow[2 * q], p.x, s2);
                s2 = fma2(arow[2 * q + 1], p.y, s2);
                if (SYM) {
                    const ulonglong2 pl = *reinterpret_cast<const ulonglong2*>(lt + c * TP + 4 * q);
                    s2 = fma2(brow[2 * q], pl.x, s2);
                    s2 = fma2(brow[2 * q + 1], pl.y, s2);
                }
            }
            float slo, shi;
            unpack2(s2, slo, shi);
            const float S = slo + shi;
            const f32x2 SS = pack2(S, S);

            const float* zcol = zt + c * CP;
            f32x2 zj[CP / 2];
#pragma unroll
            for (int q = 0; q < CP / 4; ++q) {
                const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(zcol + 4 * q);
                zj[2 * q] = p.x;
                zj[2 * q + 1] = p.y;
            }
            if constexpr (KP == 1 && BASE != 0) {
#pragma unroll
                for (int q = 0; q < CP / 2; ++q) {
                    const f32x2 d = sub2(r.z[q], zj[q]);
                    f32x2 kv, kz;
                    base_value_slope_k1<BASE>(d, r.c2[q], kv, kz);
                    gc2[q] = fma2(kv, SS, gc2[q]);
                    gz[q] = fma2(mul2(kz, SS), d, gz[q]);
                }
            } else if constexpr (KP == 1) {
#pragma unroll
                for (int q = 0; q < CP / 2; ++q) {
                    const f32x2 d = sub2(r.z[q], zj[q]);
                    const f32x2 u = fma2(d, d, r.c2[q]);
                    float ul, uh;
                    unpack2(u, ul, uh);
                    const f32x2 w = mul2(pack2(ex2_ftz(-ul), ex2_ftz(-uh)), SS);
                    gc2[q] = add2(gc2[q], w);
                    gz[q] = fma2(w, d, gz[q]);
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    f32x2 d[KP / 2];
                    f32x2 sq = pack2(BASE == 0 ? r.cg[g] : 0.f, 0.f);
#pragma unroll
                    for (int p = 0; p < KP / 2; ++p) {
                        d[p] = sub2(r.z[g * (KP / 2) + p], zj[g * (KP / 2) + p]);
                        sq = fma2(d[p], d[p], sq);
                    }
                    float lo, hi, kv, kz;
                    unpack2(sq, lo, hi);
                    base_value_slope<BASE>(lo + hi, 0.f, r.cw[g], kv, kz);
                    gcg[g] += kv * S;
                    const float w = kz * S;
                    const f32x2 ww = pack2(w, w);
#pragma unroll
                    for (int p = 0; p < KP / 2; ++p) gz[g * (KP / 2) + p] = fma2(ww, d[p], gz[g * (KP / 2) + p]);
                }
            }
        }
        __syncthreads();
        if (tid == 0 && tile + NSTAGE < ntiles) issue(tile + NSTAGE);
    }

    // ---- write dz partial (valid rows), then CTA-reduce gc ------------------------------------------------------
    if (valid) {
        const long long part = (long long)blockIdx.y * a.nchunks + chunk;
        float2* o = reinterpret_cast<float2*>(a.dz_partial + (part * a.m + row) * CP);
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) {
            float lo, hi;
            unpack2(gz[q], lo, hi);
            o[q] = make_float2(lo, hi);
        }
    }
    float gflat[GP];
    if constexpr (KP == 1) {
#pragma unroll
        for (int q = 0; q < CP / 2; ++q) unpack2(gc2[q], gflat[2 * q], gflat[2 * q + 1]);
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) gflat[g] = gcg[g];
    }
    const int lane = tid & 31, warp = tid >> 5;
    static_assert(GP <= 32, "group partials fit one warp row");
#pragma unroll
    for (int g = 0; g < GP; ++g) {
        float v = gflat[g];  // invalid rows hold zeros (their coordinates/vectors were zeroed => S == 0)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp * 32 + g] = v;
    }
    __syncthreads();
    if (tid < GP) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < ROWS_PER_CTA / 32; ++w) v += red[w * 32 + tid];
        const long long cta = (long long)blockIdx.x * a.nsplits + blockIdx.y;
        a.g_partial[(cta * a.nchunks + chunk) * GP + tid] = v;
    }
}

#endif  // __CUDACC__
}  // namespace rpgp

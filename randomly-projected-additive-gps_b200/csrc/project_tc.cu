// project_tc.cu -- the random projection Z^ = (X / l) W^T (or (X W^T) / l) on the tensor cores, and its vector-Jacobian product.
//
// Replaces `self.projection_module(x.div(self.lengthscale))` of gp_models/kernels/scaled_projection_kernel.py:21-37 and `_project`
// of polynomial_projection_kernels.py:115-137 (an nn.Linear = cuBLAS GEMM in the reference).  The product is HBM-bound (reads 4 n d
// bytes of X, writes 4 n CP nchunks of packed coordinates; the J K x d weight matrix lives in shared memory), so the kernel is
// organised around the copy: persistent CTAs walk over 128-row tiles of X,
//   * 16 producer warps read the tile in 32-column blocks (one 128-byte row segment per warp instruction, coalesced, two blocks in
//     flight per warp), split every
//     value into its tf32 part and the tf32 remainder (round-to-nearest both) and store the two operand images K-major / SWIZZLE_128B
//     into a 3-deep ring of shared-memory stages;
//   * one warp issues the tcgen05.mma kind::tf32 (split precision, all four terms: Xh.[Wh ; Wl] and Xl.[Wh ; Wl], two MMAs of N = 2 JKp per k-step) into a
//     double-buffered FP32 accumulator in tensor memory; W'' = scale * post_inv * W * pre_inv is folded, split and laid out once per CTA;
//   * 4 epilogue warps (one per TMEM lane quadrant, thread = row) read the accumulator, add the two partial sums, and write the packed
//     planes [nchunks][n][CP] (and, optionally, the natural n x JK matrix the autograd graph of the kernel classes carries) through a
//     small shared-memory transpose so that every global store is a full 128-byte line.
// Accuracy: the products of the parts are exact in FP32, the split residuals are ~2^-23 relative -- an FP32 GEMM.
// Limits of this path: d <= 128, J K <= 112; anything else takes the FP64-accumulating SIMT kernel of aux_kernels.cu.
//
// Backward (rpgp_project_bwd_f32): dW''[q][k] = sum_i dZ[i][q] X[i][k], a tall-skinny reduction over the rows with the same traffic;
// SIMT with 4 x 4 register tiles, per-CTA partials reduced in FP64 in a fixed order (bit-reproducible).
#include <algorithm>
#include <cstring>

#include "aux_kernels.cuh"
#include "sym_tc_dev.cuh"

namespace rpgp {

namespace {

using namespace tcdev;

constexpr int P_ROWS = 128;       // rows per tile
constexpr int P_NST = 3;          // A-operand stages (one 32-column block of the tile each: tf32 parts 16 KB + remainders 16 KB); 2 when 3 do not fit
constexpr int P_PROD = 16;        // producer warps (8 rows of a tile each)
constexpr int P_THREADS = 32 * (P_PROD + 1 + 4);
constexpr int P_MAX_D = 128, P_MAX_JK = 112;
#ifndef P_FENCE_PRODUCER
#define P_FENCE_PRODUCER 0
#endif
#ifndef P_DIAG
#define P_DIAG 0                  // diagnostic builds only: 1 no global loads, 2 no conversion / stores of the operand images, 4 no output stores, 8 no MMAs
#endif

// barriers
constexpr int PB_AFULL = 0;       // [3] producers -> issuer (count P_PROD)
constexpr int PB_AEMPTY = 3;      // [3] tcgen05.commit: the stage's MMAs have completed
constexpr int PB_DFULL = 6;       // [2] tcgen05.commit: a tile's accumulator is complete
constexpr int PB_DEMPTY = 8;      // [2] epilogue warps have read it (count 4)

__device__ __forceinline__ float tf32_rn_p(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
        "%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
}

struct ProjArgs {
    const float* X;
    const float* W;
    const float* pre_inv;
    const float* post_inv;
    float* Zp;          // packed planes or NULL
    float* Zn;          // natural n x JK (row stride ldz, un-scaled by `scale_nat`) or NULL
    long long n, ldx, ldz;
    int d, JK, JKp, nkb, nst;
    float scale;        // multiplies the packed output (sqrt(log2(e)/2)); the natural output carries scale_nat (1)
    Layout lay;
};

// shared memory: [B operand: nkb x (2 JKp rows x 128 B)] [A stages: P_NST x 32 KB] [staging: 4 warps x 32 x 33 floats]
//                [index tables of the epilogue: packed nchunks x 1024 + natural nchunks x 1024 uint16] [barriers]
__host__ __device__ inline uint32_t p_b_bytes(int JKp, int nkb) { return (uint32_t)nkb * 2u * (uint32_t)JKp * 128u; }
__host__ __device__ inline uint32_t p_tbl_bytes(int nchunks) { return (uint32_t)nchunks * 2u * 1024u * 2u; }
__host__ __device__ inline uint32_t p_smem_bytes(int JKp, int nkb, int nchunks, int nst) {
    return p_b_bytes(JKp, nkb) + (uint32_t)nst * 32768u + 4u * 32u * 33u * 4u + p_tbl_bytes(nchunks) + 256u + 1024u;
}
inline int p_stages(int JKp, int nkb, int nchunks) {
    for (int nst = P_NST; nst >= 2; --nst)
        if (p_smem_bytes(JKp, nkb, nchunks, nst) <= 227u * 1024u) return nst;
    return 0;
}

}  // namespace

__global__ void __launch_bounds__(P_THREADS, 1) project_tc_kernel(const ProjArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t OFF_A = p_b_bytes(a.JKp, a.nkb), OFF_STG = OFF_A + (uint32_t)a.nst * 32768u, OFF_TBL = OFF_STG + 4u * 32u * 33u * 4u;
    const uint32_t OFF_BAR = OFF_TBL + p_tbl_bytes(a.lay.nchunks);
    // epilogue index tables, built once (the integer divisions of the flattened (row, position) index cost ~300 clk per store otherwise):
    // entry = 64 * row + source column of the staged chunk (63: a padding position, stored as zero)
    uint16_t* tblp = reinterpret_cast<uint16_t*>(sm + OFF_TBL);
    uint16_t* tbln = tblp + a.lay.nchunks * 1024;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = a.d, JK = a.JK, JKp = a.JKp, nkb = a.nkb;
    const long long ntiles = (a.n + P_ROWS - 1) / P_ROWS;

    // ---- once per CTA: W'' = scale_w * post_inv[q] * W[q][k] * pre_inv[k], split, K-major SWIZZLE_128B, [tf32 parts ; remainders] stacked on N
    for (int e = tid; e < nkb * 32 * JKp; e += P_THREADS) {
        const int q = e / (nkb * 32), k = e - q * (nkb * 32);
        float w = 0.f;
        if (q < JK && k < d) {
            double v = (double)__ldg(a.W + (long long)q * d + k);
            if (a.pre_inv) v *= (double)__ldg(a.pre_inv + k);
            if (a.post_inv) v *= (double)__ldg(a.post_inv + q);
            w = (float)v;
        }
        const float h = tf32_rn_p(w), l = tf32_rn_p(w - h);
        const int kb = k >> 5, kk = k & 31;
        unsigned char* blk = sm + (size_t)kb * (2u * JKp * 128u);
        *reinterpret_cast<float*>(blk + (uint32_t)(q >> 3) * 1024u + sw128_5((uint32_t)(q & 7), (uint32_t)kk)) = h;
        const int ql = JKp + q;
        *reinterpret_cast<float*>(blk + (uint32_t)(ql >> 3) * 1024u + sw128_5((uint32_t)(ql & 7), (uint32_t)kk)) = l;
    }
    {
        const int CP = a.lay.CP, KP = a.lay.KP, G = a.lay.G, K = a.lay.K, J = a.lay.J;
        for (int e = tid; e < a.lay.nchunks * 1024; e += P_THREADS) {
            const int c = e >> 10, idx = e & 1023;
            const int q0 = c * G * K, cnt = min(G * K, JK - q0);
            uint16_t vp = 0xffff, vn = 0xffff;
            if (idx < 32 * CP) {
                const int r = idx / CP, pos = idx - r * CP, g = pos / KP, mm = pos - g * KP;
                const bool live = g < G && c * G + g < J && mm < K;
                vp = (uint16_t)(64 * r + (live ? g * K + mm : 63));
            }
            if (cnt > 0 && idx < 32 * cnt) {
                const int r = idx / cnt;
                vn = (uint16_t)(64 * r + (idx - r * cnt));
            }
            tblp[e] = vp;
            tbln[e] = vn;
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < P_NST; ++s) {
            mbar_init(&bars[PB_AFULL + s], P_PROD);
            mbar_init(&bars[PB_AEMPTY + s], 1);
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars[PB_DFULL + b], 1);
            mbar_init(&bars[PB_DEMPTY + b], 4);
        }
        mbar_fence_init();
    }
    if (warp == P_PROD) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence5_async_smem();
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < P_PROD) {
        // =========================================== producers: X tile -> split operand images ==================================
        // The kernel is issue-bound before it is HBM-bound (ncu, profiles/ncu_r02_project_cfg4_summary.md: ALU pipe 53 %, issue 68 %), so the
        // producer loop is written for instruction count: 32-bit counters advanced incrementally (no 64-bit division per block), one
        // 64-bit pointer per block, shared-memory offsets that are compile-time functions of the row, guards only where they can fail.
        constexpr int RW = P_ROWS / P_PROD;      // rows per warp
        constexpr int P_PF = 4;                  // blocks of loads in flight per warp (a ring of register sets, the loop unrolled over it)
        const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
        const int nblk = my_tiles * nkb;         // (tile, k-block) items of this CTA
        // load cursor (runs P_PF - 1 items ahead of the store cursor)
        int l_kb = 0;
        long long l_tile = blockIdx.x;
        auto load_next = [&](float* x) {
            const int col = l_kb * 32 + lane;
            const long long row0 = l_tile * P_ROWS + warp * RW;
            const float* p = a.X + row0 * a.ldx + col;
            if (P_DIAG & 1) {
#pragma unroll
                for (int r = 0; r < RW; ++r) x[r] = 0.f;
            } else if (col < d && row0 + RW <= a.n) {
#pragma unroll
                for (int r = 0; r < RW; ++r) x[r] = __ldg(p + r * a.ldx);
            } else {
#pragma unroll
                for (int r = 0; r < RW; ++r) x[r] = (col < d && row0 + r < a.n) ? __ldg(p + r * a.ldx) : 0.f;
            }
            if (++l_kb == nkb) { l_kb = 0; l_tile += gridDim.x; }
        };
        const uint32_t st_base = (uint32_t)warp * 1024u + (uint32_t)(lane & 3) * 4u;      // rows 8 warp .. 8 warp + 7 = one swizzle atom
        const uint32_t unit = (uint32_t)lane >> 2;
        float x[P_PF][RW];
#pragma unroll
        for (int u = 0; u < P_PF - 1; ++u)
            if (u < nblk) load_next(x[u]);
        int st = 0, use = 0;                     // stage of the store cursor and how often it has been used
        for (int it0 = 0; it0 < nblk; it0 += P_PF) {
#pragma unroll
            for (int u = 0; u < P_PF; ++u) {
                const int it = it0 + u;
                if (it < nblk) {
                    if (it + P_PF - 1 < nblk) load_next(x[(u + P_PF - 1) % P_PF]);
                    if (use >= 1) mbar_wait(&bars[PB_AEMPTY + st], (uint32_t)((use - 1) & 1));
                    unsigned char* ah = sm + OFF_A + (uint32_t)st * 32768u + st_base;
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        const float h = tf32_rn_p(x[u][r]), l = tf32_rn_p(x[u][r] - h);
                        const uint32_t off = (uint32_t)r * 128u + ((unit ^ (uint32_t)r) << 4);
                        if (!(P_DIAG & 2) || x[u][r] == 12345.678f) {
                            *reinterpret_cast<float*>(ah + off) = h;
                            *reinterpret_cast<float*>(ah + 16384u + off) = l;
                        }
                    }
                    // no generic -> async proxy fence here: it would wait for the prefetched global loads (MEMBAR.ALL.CTA) and serialise
                    // the copy.  The stores are released by the mbarrier arrive; the issuing warp acquires AFULL and fences the proxies
                    // before its MMAs read the stage (P_FENCE_PRODUCER=1 restores the producer-side fence).
                    if (P_FENCE_PRODUCER) fence5_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar5_arrive(&bars[PB_AFULL + st]);
                    if (++st == a.nst) { st = 0; ++use; }
                }
            }
        }
    } else if (warp == P_PROD) {
        // =========================================== MMA issue ===================================================================
        const uint32_t idesc1 = idesc5_tf32(128, 2 * JKp, 0, 0);
        const int ksteps = (d + 7) / 8;
        int t = 0, st = 0, use = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
            const int ab = t & 1;
            if (t >= 2) mbar5_wait_sleep<32>(&bars[PB_DEMPTY + ab], (uint32_t)(((t >> 1) - 1) & 1));
            tc5_fence_after();
            const uint32_t dacc = tmem + (uint32_t)(ab * 2 * JKp);
            for (int kb = 0; kb < nkb; ++kb) {
                mbar5_wait_sleep<32>(&bars[PB_AFULL + st], (uint32_t)(use & 1));
                if (!P_FENCE_PRODUCER) fence5_async_smem();
                tc5_fence_after();
                if (elect_one()) {
                    const uint32_t abase = base + OFF_A + (uint32_t)st * 32768u, bbase = base + (uint32_t)kb * (2u * JKp * 128u);
                    const int nks = min(4, ksteps - kb * 4);
#pragma unroll 1
                    for (int ks = 0; ks < ((P_DIAG & 8) ? 0 : nks); ++ks) {
                        const uint64_t dAh = smem_desc5(abase + ks * 32, 16, 1024, LAYOUT5_SW128);
                        const uint64_t dAl = smem_desc5(abase + 16384u + ks * 32, 16, 1024, LAYOUT5_SW128);
                        const uint64_t dB = smem_desc5(bbase + ks * 32, 16, 1024, LAYOUT5_SW128);
                        umma5(dacc, dAh, dB, idesc1, (kb > 0 || ks > 0) ? 1u : 0u);      // [Xh.Wh | Xh.Wl]
                        umma5(dacc, dAl, dB, idesc1, 1u);                                // [Xl.Wh | Xl.Wl]  (the fourth term comes for free)
                    }
                    umma5_commit(&bars[PB_AEMPTY + st]);
                    if (kb == nkb - 1) umma5_commit(&bars[PB_DFULL + ab]);
                }
                __syncwarp();
                if (++st == a.nst) { st = 0; ++use; }
            }
        }
    } else {
        // =========================================== epilogue: accumulator -> packed planes / natural rows ========================
        const int qd = warp & 3;                       // TMEM lane quadrant of this warp
        float* stg = reinterpret_cast<float*>(sm + OFF_STG) + qd * (32 * 33);
        const int CP = a.lay.CP, G = a.lay.G, K = a.lay.K, nchunks = a.lay.nchunks;
        long long t = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
            const int ab = (int)(t & 1);
            mbar_wait(&bars[PB_DFULL + ab], (uint32_t)((t >> 1) & 1));
            tc5_fence_after();
            const long long rowq = tile * P_ROWS + qd * 32;          // first row of this warp's quadrant
            const uint32_t tacc = tmem + (uint32_t)(ab * 2 * JKp) + ((uint32_t)(qd * 32) << 16);
            for (int c = 0; c < nchunks; ++c) {
                const int q0 = c * G * K, cnt = min(G * K, JK - q0);          // the chunk's projections q0 .. q0 + cnt - 1
                float hi[32], lo[32];
                tmem_ld32(tacc + (uint32_t)q0, hi);
                tmem_ld32(tacc + (uint32_t)(JKp + q0), lo);
#pragma unroll
                for (int e = 0; e < 32; ++e) stg[lane * 33 + e] = hi[e] + lo[e];
                __syncwarp();
                if (a.Zp) {      // 32 rows x CP floats of plane c: one contiguous range
                    float* dst = a.Zp + ((long long)c * a.n + rowq) * CP;
                    const long long lim = (a.n - rowq) * CP;           // elements of this quadrant that exist
                    const uint16_t* tb = tblp + c * 1024;
                    for (int idx = lane; idx < 32 * CP; idx += 32) {
                        const uint32_t tv = tb[idx], src = tv & 63u;
                        const float v = src != 63u ? stg[(tv >> 6) * 33u + src] * a.scale : 0.f;
                        if (idx < lim && (!(P_DIAG & 4) || v == 12345.678f)) dst[idx] = v;
                    }
                }
                if (a.Zn && cnt > 0) {
                    const uint16_t* tb = tbln + c * 1024;
                    for (int idx = lane; idx < 32 * cnt; idx += 32) {
                        const uint32_t tv = tb[idx], r = tv >> 6, e = tv & 63u;
                        if (rowq + r < a.n) a.Zn[(rowq + r) * a.ldz + q0 + e] = stg[r * 33u + e];
                    }
                }
                __syncwarp();
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar5_arrive(&bars[PB_DEMPTY + ab]);
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == P_PROD) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

bool project_tc_supported(int d, const Layout& lay) {
    const int JK = lay.J * lay.K;
    // every chunk's projections must be readable as one 32-column TMEM load inside the accumulator buffer
    const int JKp = ((JK + 15) / 16) * 16, nkb = (d + 31) / 32;
    return d >= 1 && d <= P_MAX_D && JK <= P_MAX_JK && lay.G * lay.K <= 32 && p_stages(JKp, nkb, lay.nchunks) >= 2;
}

int launch_project_tc(const float* X, long long n, int d, long long ldx, const float* W, const float* pre_inv, const float* post_inv,
                      const Layout& lay, float scale, float* Zp, float* Zn, long long ldz, cudaStream_t st) {
    if (n == 0) return OK;
    ProjArgs a;
    a.X = X; a.W = W; a.pre_inv = pre_inv; a.post_inv = post_inv; a.Zp = Zp; a.Zn = Zn;
    a.n = n; a.ldx = ldx; a.ldz = ldz; a.d = d; a.JK = lay.J * lay.K;
    a.JKp = ((a.JK + 15) / 16) * 16;
    a.nkb = (d + 31) / 32;
    a.scale = scale;
    a.lay = lay;
    a.nst = p_stages(a.JKp, a.nkb, lay.nchunks);
    const uint32_t smem = p_smem_bytes(a.JKp, a.nkb, lay.nchunks, a.nst);
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    cudaError_t e = cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(project_tc_kernel)");
    const long long ntiles = (n + P_ROWS - 1) / P_ROWS;
    project_tc_kernel<<<(unsigned)std::min<long long>(ntiles, sms), P_THREADS, smem, st>>>(a);
    note_launch();
    return cuda_fail(cudaGetLastError(), "project_tc_kernel");
}

// ---- backward: dW[q][k] = sum_i dZ[i][q] X[i][k] ------------------------------------------------------------------------------
namespace {
constexpr int PBW_THREADS = 256, PBW_RT = 32;      // threads per CTA, rows per shared-memory tile
constexpr int PBW_MAXT = 4;                        // 4 x 4 output tiles per thread (JK d <= 256 * 4 * 16)
}  // namespace

__global__ void __launch_bounds__(PBW_THREADS) project_bwd_kernel(const float* __restrict__ X, long long n, int d, long long ldx,
                                                                  const float* __restrict__ dZ, long long ldz, int JK, long long rows_per_cta,
                                                                  float* __restrict__ partial) {
    extern __shared__ float sh[];
    const int dq = (JK + 3) / 4, dk = (d + 3) / 4, JK4 = dq * 4, d4 = dk * 4;
    float* zs = sh;                    // [PBW_RT][JK4]
    float* xs = sh + PBW_RT * JK4;     // [PBW_RT][d4]
    const int ntile = dq * dk;
    float acc[PBW_MAXT][16];
#pragma unroll
    for (int u = 0; u < PBW_MAXT; ++u)
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[u][e] = 0.f;
    const long long r_begin = (long long)blockIdx.x * rows_per_cta, r_end = min(n, r_begin + rows_per_cta);
    for (long long r0 = r_begin; r0 < r_end; r0 += PBW_RT) {
        const int rows = (int)min((long long)PBW_RT, r_end - r0);
        for (int e = threadIdx.x; e < PBW_RT * JK4; e += PBW_THREADS) {
            const int r = e / JK4, q = e - r * JK4;
            zs[e] = (r < rows && q < JK) ? __ldg(dZ + (r0 + r) * ldz + q) : 0.f;
        }
        for (int e = threadIdx.x; e < PBW_RT * d4; e += PBW_THREADS) {
            const int r = e / d4, k = e - r * d4;
            xs[e] = (r < rows && k < d) ? __ldg(X + (r0 + r) * ldx + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PBW_MAXT; ++u) {
            const int tt = threadIdx.x + u * PBW_THREADS;
            if (tt < ntile) {
                const int tq = tt / dk, tk = tt - tq * dk;
#pragma unroll 4
                for (int r = 0; r < PBW_RT; ++r) {
                    const float4 z = *reinterpret_cast<const float4*>(zs + r * JK4 + 4 * tq);
                    const float4 x = *reinterpret_cast<const float4*>(xs + r * d4 + 4 * tk);
                    const float zv[4] = {z.x, z.y, z.z, z.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[u][4 * i + j] = fmaf(zv[i], xv[j], acc[u][4 * i + j]);
                }
            }
        }
        __syncthreads();
    }
    float* out = partial + (long long)blockIdx.x * JK * d;
#pragma unroll
    for (int u = 0; u < PBW_MAXT; ++u) {
        const int tt = threadIdx.x + u * PBW_THREADS;
        if (tt < ntile) {
            const int tq = tt / dk, tk = tt - tq * dk;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int q = 4 * tq + i, k = 4 * tk + j;
                    if (q < JK && k < d) out[(long long)q * d + k] = acc[u][4 * i + j];
                }
        }
    }
}

__global__ void project_bwd_reduce_kernel(const float* __restrict__ partial, int nparts, int total, float* __restrict__ dW) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += (double)partial[(long long)p * total + e];
    dW[e] = (float)s;
}

static int project_bwd_ctas(long long n) { return (int)std::max<long long>(1, std::min<long long>(296, (n + 255) / 256)); }

size_t project_bwd_workspace_bytes(long long n, int d, int JK) { return (size_t)project_bwd_ctas(n) * (size_t)JK * d * sizeof(float); }

bool project_bwd_supported(int d, int JK) { return ((JK + 3) / 4) * ((d + 3) / 4) <= PBW_THREADS * PBW_MAXT && (JK + 3) / 4 * 4 + (d + 3) / 4 * 4 <= 1500; }

int launch_project_bwd(const float* X, long long n, int d, long long ldx, const float* dZ, long long ldz, int JK, float* dW, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
    if (!project_bwd_supported(d, JK)) {
        set_error("project_bwd: J*K = %d, d = %d beyond the compiled register tiling (ceil(JK/4) ceil(d/4) <= %d)", JK, d, PBW_THREADS * PBW_MAXT);
        return ERR_UNSUPPORTED;
    }
    if (n == 0) return cuda_fail(cudaMemsetAsync(dW, 0, (size_t)JK * d * sizeof(float), st), "cudaMemsetAsync");
    const int nct = project_bwd_ctas(n);
    if (ws == nullptr || ws_bytes < project_bwd_workspace_bytes(n, d, JK)) {
        set_error("project_bwd: workspace %zu bytes < required %zu", ws_bytes, project_bwd_workspace_bytes(n, d, JK));
        return ERR_WORKSPACE;
    }
    const long long rows_per_cta = ((n + nct - 1) / nct + PBW_RT - 1) / PBW_RT * PBW_RT;
    const size_t smem = (size_t)PBW_RT * (((JK + 3) / 4) * 4 + ((d + 3) / 4) * 4) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(project_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(project_bwd_kernel)");
    }
    project_bwd_kernel<<<nct, PBW_THREADS, smem, st>>>(X, n, d, ldx, dZ, ldz, JK, rows_per_cta, (float*)ws);
    note_launch();
    RPGP_CUDA_OK(cudaGetLastError());
    const int total = JK * d;
    project_bwd_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float*)ws, nct, total, dW);
    note_launch();
    return cuda_fail(cudaGetLastError(), "project_bwd_reduce_kernel");
}

}  // namespace rpgp

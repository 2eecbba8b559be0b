// sym_tc3.cu -- symmetric K(Z,Z).V, warp-specialised: every kernel value is computed ONCE by the arithmetic warps, which also
// apply it to their own rows (row side, packed FFMA2 exactly as in the forward kernel); the transposed use
//     out[i',:] += sum_i k(i,i') V[i,:]                                   (column side)
// is a reduction ACROSS the row-owning threads and runs on the tensor cores as S^T . V_I  (tcgen05.mma kind::tf32,
// 3xTF32 split for FP32 accuracy, accumulators in TMEM).
//
// What the microbenchmarks dictated (profiles/umma_microbench_r01.txt):
//   * an MN-major tf32 operand must be in the SWIZZLE_128B_BASE32B layout -> S is written in that layout, row-locally;
//   * a UTCHMMA blocks its issuing WARP ~80-150 clk and a small MMA costs that much whatever N is -> the MMAs are issued by
//     dedicated warps (4 and 5), never by the arithmetic warps, and only the column side (which cannot be done in
//     registers) goes to the tensor core;
//   * column-side k-steps are 8 ROWS, i.e. rows of one arithmetic warp -> all hand-offs are warp-local mbarriers; there is no
//     CTA-wide barrier in the steady state and S is double buffered so the arithmetic never waits for the tensor core.
//
// CTA = 256 threads: warpgroup 0 = 4 arithmetic warps (thread = row, 128 rows, 200 registers via setmaxnreg),
// warpgroup 1 = warp 4/5 MMA issue + column-side epilogue (FP64 atomics), warp 6 TMA producer (z and V tiles), 56 registers.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "aux_kernels.cuh"
#include "kv_kernels.cuh"
#include "sym_tc.cuh"

namespace rpgp {

namespace {

constexpr int T3_ROWS = 128;     // rows per CTA
constexpr int T3_BN = 32;        // columns per tile
constexpr int T3_N = 16;         // padded right-hand sides
#ifndef T3_QUNROLL
#define T3_QUNROLL 1
#endif
constexpr int T3_QU = T3_QUNROLL;   // unroll factor of the 4-column groups inside a tile

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t T3_SC = 0;                  // S^T operand: buffer b at b*32768: hi [128 rows][128 B], lo 16384 B later
constexpr uint32_t T3_BC = 65536;              // V of the row block as B operand [32 = 16 hi rows + 16 lo rows][128 k]: 4 k-blocks x 4096 B
constexpr uint32_t T3_Z = 81920;               // z tiles: 2 stages x 32 x CP floats (<= 4096 B each)
constexpr uint32_t T3_V = 90112;               // V tiles: 2 stages x 32 x 16 floats (2048 B each)
constexpr uint32_t T3_BAR = 94208;             // mbarriers
constexpr uint32_t T3_SMEM_BYTES = T3_BAR + 256 + 1024;

// barrier indices
constexpr int B_ZFULL = 0;     // [2]  producer -> arithmetic warps (count 1 + tx bytes)
constexpr int B_ZEMPTY = 2;    // [2]  arithmetic warps -> producer (count 4)
constexpr int B_SFULL = 4;     // [4 warps][2 buffers]  arithmetic warp w -> issuer (count 1)
constexpr int B_TDONE = 12;    // [2]  both issuers' tcgen05.commit (count 2): S buffer reusable, D2 readable
constexpr int B_EREAD = 14;    // [2]  both issuers have read D2 of that buffer (count 2): accumulators reusable
constexpr int B_COUNT = 16;

constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t smem_desc3(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc3_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma3(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma3_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// helper warps poll with a back-off so that their spinning does not take issue slots from the arithmetic warps
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(128);
    }
}
__device__ __forceinline__ void tc3_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc3_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence3_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem3_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]);
}
__device__ __forceinline__ float tf32_hi3(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t sw128_3(uint32_t row, uint32_t kk) {
    return row * 128u + ((((kk >> 2) ^ (row & 7u)) << 4) | ((kk & 3u) << 2));
}

struct Sym3Args {
    const float* z;        // [n][CP]
    const float* v;        // [n][16]
    const float* nlc;      // [CP]
    double* acc;           // [n][16] FP64 accumulators (zeroed by the launcher)
    long long n;
    int nblocks, half, nsplits, rb_begin;
};

// tile enumeration shared by all roles: offsets k in [k_begin, k_end), four 32-column tiles per 128-column block
struct TileIter {
    int I, B, k_begin, ntiles;
    long long n;
    __device__ __forceinline__ int block_of(int k) const { int Ip = I + k; return Ip >= B ? Ip - B : Ip; }
    __device__ __forceinline__ bool offset_active(int k) const { return !((B % 2 == 0) && (k == B / 2) && (I >= B / 2)); }
    __device__ __forceinline__ long long col0(int t) const { return (long long)block_of(k_begin + (t >> 2)) * T3_ROWS + (t & 3) * T3_BN; }
    __device__ __forceinline__ bool live(int t) const { return offset_active(k_begin + (t >> 2)) && col0(t) < n; }
    __device__ __forceinline__ bool diag(int t) const { return block_of(k_begin + (t >> 2)) == I; }
    __device__ __forceinline__ int next_live(int t) const { while (t < ntiles && !live(t)) ++t; return t; }
};

}  // namespace

template <int CP, int NP2, int TP>
__global__ void __launch_bounds__(256, 2) mvm_sym_tc3_kernel(const Sym3Args a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + T3_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + T3_BAR + 192);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TileIter it;
    it.I = a.rb_begin + blockIdx.x;
    it.B = a.nblocks;
    it.n = a.n;
    const int per = (a.half + a.nsplits - 1) / a.nsplits;
    it.k_begin = blockIdx.y * per;
    it.ntiles = 4 * (min(a.half, it.k_begin + per) - it.k_begin);
    if (it.ntiles < 0) it.ntiles = 0;

    if (tid == 0) {
        mbar_init(&bars[B_ZFULL + 0], 1);
        mbar_init(&bars[B_ZFULL + 1], 1);
        mbar_init(&bars[B_ZEMPTY + 0], 4);
        mbar_init(&bars[B_ZEMPTY + 1], 4);
#pragma unroll
        for (int w = 0; w < 8; ++w) mbar_init(&bars[B_SFULL + w], 1);
        mbar_init(&bars[B_TDONE + 0], 2);
        mbar_init(&bars[B_TDONE + 1], 2);
        mbar_init(&bars[B_EREAD + 0], 2);
        mbar_init(&bars[B_EREAD + 1], 2);
        mbar_fence_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc3_fence_before();
    __syncthreads();
    tc3_fence_after();
    const uint32_t tmem = *tmem_slot;   // D2[issuer h][buffer b]: 32 columns (Sh.Vh + Sl.Vh | Sh.Vl) at column 64*b + 32*h

    if (warp < 4) {
        // =========================================== arithmetic warps ===================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const long long row = (long long)it.I * T3_ROWS + tid;
        const bool valid = row < a.n;
        RowCoords<CP, 1, CP> r;
        load_row_coords<CP, 1, CP>(r, a.z + row * CP, valid, a.nlc);
        {   // B operand of the column side: V rows of this block, split, [16 rows c][128 k = row-in-block], K-major SW128
            const int kb = tid >> 5, kk = tid & 31;
#pragma unroll
            for (int c = 0; c < T3_N; ++c) {
                const float v = valid ? __ldg(a.v + row * T3_N + c) : 0.f;
                const float h = tf32_hi3(v);
                const uint32_t off = (uint32_t)kb * 4096u + (uint32_t)(c >> 3) * 1024u + sw128_3((uint32_t)(c & 7), (uint32_t)kk);
                *reinterpret_cast<float*>(sm + T3_BC + off) = h;                 // rows 0..15: tf32-hi part
                *reinterpret_cast<float*>(sm + T3_BC + 2048u + off) = v - h;     // rows 16..31: remainder
            }
        }
        f32x2 acc[TP / 2], comp[TP / 2];   // TP = right-hand sides rounded up to 4 (row side only; the MMA N stays 16)
#pragma unroll
        for (int q = 0; q < TP / 2; ++q) { acc[q] = 0ull; comp[q] = 0ull; }

        int j = 0, jc = 0;
        for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++j) {
            const int s = j & 1;
            const long long c0 = it.col0(t);
            const int cols = (int)min((long long)T3_BN, a.n - c0);
            const bool diag = it.diag(t);
            const int bc = jc & 1;
            mbar_wait(&bars[B_ZFULL + s], (uint32_t)((j >> 1) & 1));
            if (!diag && jc >= 2) mbar_wait(&bars[B_TDONE + bc], (uint32_t)(((jc >> 1) - 1) & 1));   // S buffer bc is free again
            const float* zt = reinterpret_cast<const float*>(sm + T3_Z) + (size_t)s * T3_BN * CP;
            const float* vt = reinterpret_cast<const float*>(sm + T3_V) + (size_t)s * T3_BN * T3_N;
            unsigned char* sc = sm + T3_SC + (uint32_t)bc * 32768u;
            f32x2 lo[TP / 2];
#pragma unroll
            for (int q = 0; q < TP / 2; ++q) lo[q] = 0ull;
            auto tile_body = [&](auto full_tile) {
                constexpr bool FULL = decltype(full_tile)::value;
#pragma unroll T3_QU
                for (int q = 0; q < T3_BN / 4; ++q) {
                    float sv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 4 * q + e;
                        const float val = pair_kernel_value<CP, 1, CP, NP2>(r, zt + c * CP);
                        sv[e] = (FULL || c < cols) ? val : 0.f;      // columns behind a partial tile hold stale bytes
                        const f32x2 ss = pack2(sv[e], sv[e]);
#pragma unroll
                        for (int k = 0; k < TP / 4; ++k) {   // row side: out[i,:] += k(i,i') V[i',:]
                            const ulonglong2 p = *reinterpret_cast<const ulonglong2*>(vt + c * T3_N + 4 * k);
                            lo[2 * k] = fma2(ss, p.x, lo[2 * k]);
                            lo[2 * k + 1] = fma2(ss, p.y, lo[2 * k + 1]);
                        }
                    }
                    if (!diag) {   // column side operand: split, SWIZZLE_128B_BASE32B (32-byte chunks ^ row % 4), row-local stores
                        float4 h, l;
                        h.x = tf32_hi3(sv[0]); h.y = tf32_hi3(sv[1]); h.z = tf32_hi3(sv[2]); h.w = tf32_hi3(sv[3]);
                        l.x = sv[0] - h.x; l.y = sv[1] - h.y; l.z = sv[2] - h.z; l.w = sv[3] - h.w;
                        const uint32_t off = (uint32_t)tid * 128u + (((((uint32_t)q >> 1) ^ ((uint32_t)tid & 3u)) << 5) | (((uint32_t)q & 1u) << 4));
                        *reinterpret_cast<float4*>(sc + off) = h;
                        *reinterpret_cast<float4*>(sc + 16384u + off) = l;
                    }
                }
            };
            if (cols == T3_BN) tile_body(std::true_type{}); else tile_body(std::false_type{});
#pragma unroll
            for (int q = 0; q < TP / 2; ++q) {   // Kahan-compensated fold of the tile sum
                const f32x2 y = sub2(lo[q], comp[q]);
                const f32x2 tsum = add2(acc[q], y);
                comp[q] = sub2(sub2(tsum, acc[q]), y);
                acc[q] = tsum;
            }
            if (!diag) {
                fence3_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars[B_SFULL + warp * 2 + bc]);
                ++jc;
            } else {
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&bars[B_ZEMPTY + s]);
        }
        if (valid && j > 0) {
            double* dst = a.acc + row * T3_N;
#pragma unroll
            for (int q = 0; q < TP / 2; ++q) {
                float x, y;
                unpack2(acc[q], x, y);
                atomicAdd(dst + 2 * q, (double)x);
                atomicAdd(dst + 2 * q + 1, (double)y);
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 6) {
            // =========================================== TMA producer ====================================================
            if (lane == 0) {
                int j = 0;
                for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++j) {
                    const int s = j & 1;
                    if (j >= 2) mbar_wait_sleep(&bars[B_ZEMPTY + s], (uint32_t)(((j >> 1) - 1) & 1));
                    const long long c0 = it.col0(t);
                    const uint32_t cols = (uint32_t)min((long long)T3_BN, a.n - c0);
                    mbar_expect_tx(&bars[B_ZFULL + s], cols * (CP + T3_N) * (uint32_t)sizeof(float));
                    bulk_g2s(sm + T3_Z + (uint32_t)s * T3_BN * CP * 4u, a.z + c0 * CP, cols * CP * (uint32_t)sizeof(float), &bars[B_ZFULL + s]);
                    bulk_g2s(sm + T3_V + (uint32_t)s * T3_BN * T3_N * 4u, a.v + c0 * T3_N, cols * T3_N * (uint32_t)sizeof(float), &bars[B_ZFULL + s]);
                }
            }
        } else if (warp == 4 || warp == 5) {
            // =========================================== MMA issue + column-side epilogue =================================
            const int h = warp - 4;                      // arithmetic warps 2h, 2h+1  ->  k-groups 8h .. 8h+7; TMEM lane quadrant h
            // two MMAs per k-step instead of three: Sh.[Vh | Vl] (N = 32) and Sl.Vh (N = 16, into the first 16 columns)
            constexpr uint32_t IDESC_N32 = idesc3_tf32(64, 2 * T3_N, 1, 0), IDESC_N16 = idesc3_tf32(64, T3_N, 1, 0);
            const uint64_t dB = smem_desc3(base + T3_BC, 16, 1024, LAYOUT_SW128);
            int jc = 0;
            long long prev_c0 = -1;
            auto epilogue = [&](int pjc, long long pc0) {
                const int pb = pjc & 1;
                mbar_wait_sleep(&bars[B_TDONE + pb], (uint32_t)((pjc >> 1) & 1));
                tc3_fence_after();
                float d0[16], d1[16];
                const uint32_t lane_base = (uint32_t)(h * 32) << 16;
                {   // sum of the four 16-column pieces: two issuers x (Sh.Vh + Sl.Vh | Sh.Vl)
                    float e0[16], e1[16];
                    tmem3_ld16(tmem + 64u * pb + lane_base, d0);
                    tmem3_ld16(tmem + 64u * pb + 16u + lane_base, e0);
                    tmem3_ld16(tmem + 64u * pb + 32u + lane_base, d1);
                    tmem3_ld16(tmem + 64u * pb + 48u + lane_base, e1);
#pragma unroll
                    for (int c = 0; c < 16; ++c) { d0[c] += e0[c]; d1[c] += e1[c]; }
                }
                tc3_fence_before();
                if (lane == 0) mbar_arrive(&bars[B_EREAD + pb]);
                const long long crow = pc0 + 16 * h + lane;   // M = 64 layout: rows 16h .. 16h+15 in lanes 0..15 of quadrant h
                if (lane < 16 && crow < a.n) {
                    double* dst = a.acc + crow * T3_N;
#pragma unroll
                    for (int c = 0; c < T3_N; ++c) atomicAdd(dst + c, (double)d0[c] + (double)d1[c]);
                }
            };
            for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1)) {
                if (it.diag(t)) continue;
                const int bc = jc & 1;
                if (jc >= 2) mbar_wait_sleep(&bars[B_EREAD + bc], (uint32_t)(((jc >> 1) - 1) & 1));   // both issuers have read D2 of tile jc-2
                const uint32_t d2 = tmem + 64u * bc + 32u * h;
                const uint64_t dA_h = smem_desc3(base + T3_SC + (uint32_t)bc * 32768u, 16384, 512, LAYOUT_SW128_BASE32B);
                const uint64_t dA_l = smem_desc3(base + T3_SC + (uint32_t)bc * 32768u + 16384u, 16384, 512, LAYOUT_SW128_BASE32B);
#pragma unroll 1
                for (int sw = 2 * h; sw < 2 * h + 2; ++sw) {
                    mbar_wait_sleep(&bars[B_SFULL + sw * 2 + bc], (uint32_t)((jc >> 1) & 1));
                    tc3_fence_after();
                    if (lane == 0) {
#pragma unroll
                        for (int gl = 0; gl < 4; ++gl) {
                            const int g = 4 * sw + gl;
                            const uint64_t aoff = (uint64_t)((g * 1024) >> 4);
                            const uint64_t boff = (uint64_t)(((g >> 2) * 4096 + (g & 3) * 32) >> 4);
                            umma3(d2, dA_h + aoff, dB + boff, IDESC_N32, (sw > 2 * h || gl > 0) ? 1u : 0u);
                            umma3(d2, dA_l + aoff, dB + boff, IDESC_N16, 1);
                        }
                    }
                    __syncwarp();
                }
                if (lane == 0) umma3_commit(&bars[B_TDONE + bc]);
                __syncwarp();
                if (prev_c0 >= 0) epilogue(jc - 1, prev_c0);
                prev_c0 = it.col0(t);
                ++jc;
            }
            if (prev_c0 >= 0) epilogue(jc - 1, prev_c0);
        }
    }
    tc3_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
    }
}

template <int CP, int NP2, int TP>
static int run_sym3(const Sym3Args& a, dim3 grid, cudaStream_t st) {
    auto kernel = mvm_sym_tc3_kernel<CP, NP2, TP>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T3_SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mvm_sym_tc3_kernel)");
    kernel<<<grid, 256, T3_SMEM_BYTES, st>>>(a);
    note_launch();
    return cuda_fail(cudaGetLastError(), "mvm_sym_tc3_kernel launch");
}

__global__ void sym3_finalize_kernel(const double* __restrict__ acc, long long n, int t, float* __restrict__ out, int ldo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * t) return;
    const long long row = idx / t;
    const int c = (int)(idx - row * t);
    out[row * ldo + c] = (float)acc[row * T3_N + c];
}

int launch_sym_tc3(const float* zp, long long n, int CP, const float* nlc, const float* V16, int t, float* out, int ldo,
                   int rb_begin, int rb_end, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const size_t need = (size_t)n * T3_N * sizeof(double);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("mvm_sym: workspace %zu bytes < required %zu", workspace_bytes, need);
        return ERR_WORKSPACE;
    }
    double* acc = (double*)workspace;
    RPGP_CUDA_OK(cudaMemsetAsync(acc, 0, need, st));
    Sym3Args a;
    a.z = zp; a.v = V16; a.nlc = nlc; a.acc = acc; a.n = n;
    a.nblocks = (int)((n + T3_ROWS - 1) / T3_ROWS);
    a.half = a.nblocks / 2 + 1;
    a.rb_begin = rb_begin;
    const int nrb = rb_end - rb_begin;
    if (nrb > 0) {
        static const int splits_env = [] { const char* e = getenv("RPGP_SYM_SPLITS"); return e ? atoi(e) : 0; }();
        long long want = (148LL * 2 * 16 + nrb - 1) / nrb;
        if (splits_env > 0) want = splits_env;
        want = std::max<long long>(1, std::min<long long>(want, a.half));
        a.nsplits = (int)want;
        dim3 grid((unsigned)nrb, (unsigned)a.nsplits, 1);
        // polynomial-exp2 pairs: this kernel is bound by instruction issue (an FFMA2 takes two issue cycles), so fewer pairs
        // go to the FMA pipe than in the forward kernel; RPGP_SYM_POLY_PAIRS overrides (tools sweep)
        static const int np_env = [] { const char* e = getenv("RPGP_SYM_POLY_PAIRS"); return e ? atoi(e) : -1; }();
        const int tp = t <= 4 ? 4 : (t <= 8 ? 8 : (t <= 12 ? 12 : 16));
        int rc = ERR_UNSUPPORTED;
#define RPGP_SYM3_CASE(CPv, NPv)                                                                              \
        if (CP == CPv && np == NPv) {                                                                         \
            rc = tp == 4 ? run_sym3<CPv, NPv, 4>(a, grid, st) : tp == 8 ? run_sym3<CPv, NPv, 8>(a, grid, st)    \
                 : tp == 12 ? run_sym3<CPv, NPv, 12>(a, grid, st) : run_sym3<CPv, NPv, 16>(a, grid, st);       \
        }
        const int np = np_env >= 0 ? np_env : (CP >= 20 ? 2 : (CP >= 16 ? 1 : 0));
        RPGP_SYM3_CASE(4, 0) RPGP_SYM3_CASE(8, 0) RPGP_SYM3_CASE(12, 0) RPGP_SYM3_CASE(16, 0) RPGP_SYM3_CASE(16, 1)
        RPGP_SYM3_CASE(20, 0) RPGP_SYM3_CASE(20, 1) RPGP_SYM3_CASE(20, 2) RPGP_SYM3_CASE(24, 0) RPGP_SYM3_CASE(24, 2)
        RPGP_SYM3_CASE(28, 0) RPGP_SYM3_CASE(28, 2) RPGP_SYM3_CASE(32, 0) RPGP_SYM3_CASE(32, 2)
#undef RPGP_SYM3_CASE
        if (rc == ERR_UNSUPPORTED) set_error("mvm_sym: no kernel for CP=%d poly pairs=%d (compiled: 0 for every CP, 1 for CP 16/20, 2 for CP >= 20)", CP, np);
        if (rc) return rc;
    }
    const long long total = n * t;
    sym3_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(acc, n, t, out, ldo);
    note_launch();
    return cuda_fail(cudaGetLastError(), "sym3_finalize_kernel");
}

}  // namespace rpgp

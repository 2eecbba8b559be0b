// sym_tc.cu -- symmetric K(Z,Z).V with every kernel value computed ONCE and both of its uses,
//     out[i ,:] += k(i,i') V[i',:]     (row side)        out[i',:] += k(i,i') V[i ,:]     (column side)
// carried by the 5th-generation tensor cores (tcgen05.mma, kind::tf32, accumulators in TMEM).
//
// Why: the forward kernel of kv_kernels.cuh sits on the XU-pipe roof (one MUFU.EX2 per (i,i',j)); K(Z,Z) is symmetric,
// so half of those exponentials are redundant.  Using a value for the column side needs a reduction ACROSS the threads
// that own the rows -- exactly a matrix product S^T . V_I, which is what tensor cores are for.  FP32 accuracy is kept
// with the 3xTF32 split (S = Sh + Sl, V = Vh + Vl, Sh.Vh + Sl.Vh + Sh.Vl; every dropped term is < 2^-21 relative) and by
// flushing the TMEM accumulators into FP32 registers / FP64 global accumulators after every 64-column tile.
//
// CTA = 128 rows (one thread per row, 4 warps), column tiles of 32.  Per tile:
//   1. each thread forms s(i,i') for its row and 32 columns exactly like the forward kernel (packed FADD2/FFMA2, MUFU.EX2,
//      optional FMA-pipe polynomial pairs), splits it into Sh/Sl and writes its row of the [128 rows][32 cols] tile twice:
//      once in the 128B-swizzled K-major layout and once in the SWIZZLE_128B_BASE32B layout.  (kind::tf32 accepts an
//      MN-major operand only with the 32-byte-base swizzle and a K-major operand only with the 16-byte-base swizzles --
//      measured with tools/umma_probe*.cu -- so one set of bytes cannot serve both products; both copies are plain
//      16-byte row-local stores.)
//   2. the tensor core reads the first copy as the K-major A operand (M=128 rows, K=32 columns) of the row side and the
//      second as the MN-major A operand (M=64: 32 columns + 32 ignored, K=128 rows) of the column side; B operands are
//      the split V tiles ([16][K], K-major);
//   3. tcgen05.commit -> mbarrier; the next tile's arithmetic of the co-resident CTA hides the MMA latency; D is read
//      back with tcgen05.ld, row side into Kahan-compensated registers, column side as FP64 atomics.
// Unique block pairs {I, I'} are enumerated cyclically (I' = I + k mod B, k <= B/2) so every CTA has the same work.
#include <algorithm>

#include "aux_kernels.cuh"
#include "kv_kernels.cuh"
#include "sym_tc.cuh"

namespace rpgp {

namespace {

constexpr int SYM_BM = 128;      // rows per CTA (= threads)
constexpr int SYM_BN = 32;       // columns per tile
constexpr int SYM_N = 16;        // padded right-hand sides (MMA N)

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t OFF_SRH = 0;                // S hi, row-side copy   [128 rows][128 B], SWIZZLE_128B (K-major A)
constexpr uint32_t OFF_SRL = 16384;            // S lo, row-side copy
constexpr uint32_t OFF_SCH = 32768;            // S hi, column-side copy [128 rows][128 B], SWIZZLE_128B_BASE32B (MN-major A)
constexpr uint32_t OFF_SCL = 49152;            // S lo, column-side copy  (the block 16 KB after each copy is what the
                                               //  M=64 descriptor sees as columns 32..63: ignored on read-back)
constexpr uint32_t OFF_BCH = 65536;            // column-side B (V of the row block, [16][128 rows]) hi: 4 k-blocks x 2048 B
constexpr uint32_t OFF_BCL = 73728;            //                                                     lo
constexpr uint32_t OFF_BRH = 81920;            // row-side B (V^T of the column tile, [16][32 cols]) hi: 2048 B
constexpr uint32_t OFF_BRL = 83968;            //                                                    lo
constexpr uint32_t OFF_Z = 86016;              // z tiles: 2 stages x 32 x CP floats (<= 4096 B each)
constexpr uint32_t OFF_BAR = 94208;            // mbarriers: full[2], mma_done; tmem base slot
constexpr uint32_t SYM_SMEM_BYTES = OFF_BAR + 64 + 1024;  // + alignment slack

constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = LAYOUT_SW128) {
    // UMMA shared-memory descriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout type [61,64)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor, kind::tf32, FP32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
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

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// byte offset of element (row, kk) inside one [rows][128 B] swizzled block (kk = 0..31 along the 128-byte line)
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t kk) {
    return row * 128u + ((((kk >> 2) ^ (row & 7u)) << 4) | ((kk & 3u) << 2));
}

}  // namespace

struct SymArgs {
    const float* z;        // [n][CP] packed coordinates (single chunk)
    const float* v;        // [n][16] right-hand sides, zero padded
    const float* vht;      // [16][npad] tf32-hi part of V, transposed, zero padded
    const float* vlt;      // [16][npad] remainder
    const float* nlc;      // [CP]
    double* acc;           // [n][16] FP64 accumulators (zeroed by the caller)
    long long n, npad;
    int nblocks;           // B = ceil(n / 128)
    int half;              // floor(B/2) + 1 column-block offsets per row block
    int nsplits;           // column-offset splits per row block
    int rb_begin, rb_end;  // row blocks handled by this launch (rank partition)
};

template <int CP, int NP2>
__global__ void __launch_bounds__(SYM_BM, 2) mvm_sym_tc_kernel(const SymArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + OFF_BAR);   // [0],[1]: z tile full; [2]: mma done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 32);
    float* ztile = reinterpret_cast<float*>(sm + OFF_Z);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int I = a.rb_begin + blockIdx.x;
    const int B = a.nblocks;
    // column-block offsets handled by this CTA
    const int per = (a.half + a.nsplits - 1) / a.nsplits;
    const int k_begin = blockIdx.y * per;
    const int k_end = min(a.half, k_begin + per);
    const long long row = (long long)I * SYM_BM + tid;
    const bool valid = row < a.n;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_d1 = tmem;                  // row-side accumulator  [128 lanes][16 cols]
    const uint32_t t_d2 = tmem + 16;             // column-side accumulator (M = 64 layout: rows 16w..16w+15 in lanes 0..15 of warp w)

    // ---- per-row state ---------------------------------------------------------------------------------------------
    RowCoords<CP, 1, CP> r;
    load_row_coords<CP, 1, CP>(r, a.z + row * CP, valid, a.nlc);
    {   // column-side B operand: V rows of this block, split, as [16 rows c][128 k = row-in-block], K-major SW128
        const int kb = tid >> 5;               // 32 rows per 128-byte k-block
        const int kk = tid & 31;
#pragma unroll
        for (int c = 0; c < SYM_N; ++c) {
            const float v = valid ? __ldg(a.v + row * SYM_N + c) : 0.f;
            const float h = tf32_hi(v);
            const uint32_t off = (uint32_t)kb * 2048u + (uint32_t)(c >> 3) * 1024u + sw128((uint32_t)(c & 7), (uint32_t)kk);
            *reinterpret_cast<float*>(sm + OFF_BCH + off) = h;
            *reinterpret_cast<float*>(sm + OFF_BCL + off) = v - h;
        }
    }
    float acc[SYM_N], comp[SYM_N];
#pragma unroll
    for (int c = 0; c < SYM_N; ++c) { acc[c] = 0.f; comp[c] = 0.f; }

    // enumerate the tiles of this CTA: offsets k in [k_begin, k_end), four 32-column tiles per 128-column block
    auto block_of = [&](int k) { int Ip = I + k; return Ip >= B ? Ip - B : Ip; };
    auto offset_active = [&](int k) {  // for even B the antipodal offset is shared by two row blocks: the lower one takes it
        return !((B % 2 == 0) && (k == B / 2) && (I >= B / 2));
    };
    const int ntiles = 4 * (k_end - k_begin);
    auto tile_col0 = [&](int t) { return (long long)block_of(k_begin + (t >> 2)) * SYM_BM + (t & 3) * SYM_BN; };
    auto tile_live = [&](int t) { return offset_active(k_begin + (t >> 2)) && tile_col0(t) < a.n; };

    int issued = 0;  // live tiles whose z rows have been requested (thread 0 only)
    auto issue_z = [&](int t) {
        const int s = issued & 1;
        ++issued;
        const long long c0 = tile_col0(t);
        const uint32_t cols = (uint32_t)max(0ll, min((long long)SYM_BN, a.n - c0));
        mbar_expect_tx(&bars[s], cols * CP * (uint32_t)sizeof(float));
        bulk_g2s(ztile + (size_t)s * SYM_BN * CP, a.z + c0 * CP, cols * CP * (uint32_t)sizeof(float), &bars[s]);
    };
    // first live tiles
    int next_issue = 0;
    auto advance_issue = [&]() {
        while (next_issue < ntiles && !tile_live(next_issue)) ++next_issue;
    };
    uint32_t zphase[2] = {0, 0};
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            advance_issue();
            if (next_issue < ntiles) { issue_z(next_issue); }
            ++next_issue;
        }
    }

    constexpr uint32_t IDESC_ROW = idesc_tf32(128, SYM_N, 0, 0);
    constexpr uint32_t IDESC_COL = idesc_tf32(64, SYM_N, 1, 0);
    uint32_t mma_phase = 0;
    int prev_tile = -1;         // tile whose MMAs are in flight
    bool prev_col_side = false;
    long long prev_col0 = 0;
    int nlive = 0;

    auto drain_prev = [&]() {
        // wait for the MMAs of the previous tile, fold D1 into the register accumulators, push D2 to global memory
        mbar_wait(&bars[2], mma_phase);
        mma_phase ^= 1u;
        tc_fence_after();
        float d[16];
        tmem_ld16(t_d1 + ((uint32_t)(warp * 32) << 16), d);
#pragma unroll
        for (int c = 0; c < SYM_N; ++c) {  // Kahan-compensated fold
            const float y = d[c] - comp[c];
            const float tsum = acc[c] + y;
            comp[c] = (tsum - acc[c]) - y;
            acc[c] = tsum;
        }
        if (prev_col_side && warp < 2) {           // 32 live columns: D2 rows 0..31 sit in warps 0 and 1
            tmem_ld16(t_d2 + ((uint32_t)(warp * 32) << 16), d);
            const long long crow = prev_col0 + warp * 16 + lane;   // M = 64 layout: lanes 0..15 of warp w hold rows 16w..16w+15
            if (lane < 16 && crow < a.n) {
                double* dst = a.acc + crow * SYM_N;
#pragma unroll
                for (int c = 0; c < SYM_N; ++c) atomicAdd(dst + c, (double)d[c]);
            }
        }
        tc_fence_before();
    };

    int zi = 0;  // index among live tiles -> z stage
    for (int t = 0; t < ntiles; ++t) {
        if (!tile_live(t)) continue;
        const int s = zi & 1;
        const long long c0 = tile_col0(t);
        const int cols = (int)min((long long)SYM_BN, a.n - c0);
        const bool diag = (block_of(k_begin + (t >> 2)) == I);
        // (1) previous tile's MMAs must be done before S / B_row are overwritten
        if (prev_tile >= 0) drain_prev();
        // (2) row-side B operand: V^T tile, split planes, [16 rows c][32 cols] = one k-block of [16][128 B]
        {
            const int c = tid >> 3, q = tid & 7;           // q: 16-byte chunk (4 columns) of the 32-column tile
            const long long col = c0 + 4 * q;
            const float4 h = *reinterpret_cast<const float4*>(a.vht + (long long)c * a.npad + col);
            const float4 l = *reinterpret_cast<const float4*>(a.vlt + (long long)c * a.npad + col);
            const uint32_t off = (uint32_t)(c >> 3) * 1024u + (uint32_t)(c & 7) * 128u + ((((uint32_t)q) ^ ((uint32_t)c & 7u)) << 4);
            *reinterpret_cast<float4*>(sm + OFF_BRH + off) = h;
            *reinterpret_cast<float4*>(sm + OFF_BRL + off) = l;
        }
        // (3) kernel values of this tile
        mbar_wait(&bars[s], zphase[s]);
        zphase[s] ^= 1u;
        const float* zt = ztile + (size_t)s * SYM_BN * CP;
#pragma unroll 1
        for (int q = 0; q < SYM_BN / 4; ++q) {
            float sv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = 4 * q + e;
                float val = pair_kernel_value<CP, 1, CP, NP2>(r, zt + c * CP);
                sv[e] = (c < cols) ? val : 0.f;          // stale shared memory behind a partial tile must not leak NaNs
            }
            float4 h, l;
            h.x = tf32_hi(sv[0]); h.y = tf32_hi(sv[1]); h.z = tf32_hi(sv[2]); h.w = tf32_hi(sv[3]);
            l.x = sv[0] - h.x; l.y = sv[1] - h.y; l.z = sv[2] - h.z; l.w = sv[3] - h.w;
            const uint32_t off_r = (uint32_t)tid * 128u + ((((uint32_t)q) ^ ((uint32_t)tid & 7u)) << 4);          // 16 B chunks ^ row%8
            const uint32_t off_c = (uint32_t)tid * 128u + (((((uint32_t)q >> 1) ^ ((uint32_t)tid & 3u)) << 5) | (((uint32_t)q & 1u) << 4));  // 32 B chunks ^ row%4
            *reinterpret_cast<float4*>(sm + OFF_SRH + off_r) = h;
            *reinterpret_cast<float4*>(sm + OFF_SRL + off_r) = l;
            *reinterpret_cast<float4*>(sm + OFF_SCH + off_c) = h;
            *reinterpret_cast<float4*>(sm + OFF_SCL + off_c) = l;
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        // (4) refill the z stage, issue the MMAs
        if (tid == 0) {
            advance_issue();
            if (next_issue < ntiles) issue_z(next_issue);
            ++next_issue;
            tc_fence_after();
            // row side: D1[128 x 16] = S[128 x 32] . V_tile[32 x 16]   (A K-major SWIZZLE_128B, 4 k-steps of 8 columns)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t koff = (uint32_t)ks * 32u;
                const uint64_t ah = smem_desc(base + OFF_SRH + koff, 16, 1024), al = smem_desc(base + OFF_SRL + koff, 16, 1024);
                const uint64_t bh = smem_desc(base + OFF_BRH + koff, 16, 1024), bl = smem_desc(base + OFF_BRL + koff, 16, 1024);
                umma_tf32(t_d1, ah, bh, IDESC_ROW, ks > 0);
                umma_tf32(t_d1, al, bh, IDESC_ROW, 1);
                umma_tf32(t_d1, ah, bl, IDESC_ROW, 1);
            }
            // column side: D2[64 x 16] = S^T[64 x 128] . V_I[128 x 16]   (A MN-major SWIZZLE_128B_BASE32B: 4-row atoms of
            // 512 B, two per k-step of 8 rows; rows 32..63 of D2 come from the neighbouring 16 KB and are never read)
            if (!diag) {
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    const uint32_t aoff = (uint32_t)g * 1024u;
                    const uint32_t boff = (uint32_t)(g >> 2) * 2048u + (uint32_t)(g & 3) * 32u;
                    const uint64_t ah = smem_desc(base + OFF_SCH + aoff, 16384, 512, LAYOUT_SW128_BASE32B);
                    const uint64_t al = smem_desc(base + OFF_SCL + aoff, 16384, 512, LAYOUT_SW128_BASE32B);
                    const uint64_t bh = smem_desc(base + OFF_BCH + boff, 16, 1024), bl = smem_desc(base + OFF_BCL + boff, 16, 1024);
                    umma_tf32(t_d2, ah, bh, IDESC_COL, g > 0);
                    umma_tf32(t_d2, al, bh, IDESC_COL, 1);
                    umma_tf32(t_d2, ah, bl, IDESC_COL, 1);
                }
            }
            umma_commit(&bars[2]);
        }
        prev_tile = t;
        prev_col_side = !diag;
        prev_col0 = c0;
        ++nlive;
        ++zi;
    }
    if (prev_tile >= 0) drain_prev();

    // ---- row-side result of this CTA -> global FP64 accumulators -----------------------------------------------------
    if (valid && nlive > 0) {
        double* dst = a.acc + row * SYM_N;
#pragma unroll
        for (int c = 0; c < SYM_N; ++c) atomicAdd(dst + c, (double)acc[c]);
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
    }
}

// V (n x ldv) -> split, transposed planes [16][npad]
__global__ void sym_prep_v_kernel(const float* __restrict__ V, long long n, int ldv, int t, long long npad,
                                  float* __restrict__ vht, float* __restrict__ vlt) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= npad * SYM_N) return;
    const int c = (int)(idx / npad);
    const long long i = idx - (long long)c * npad;
    const float v = (i < n && c < t) ? V[i * ldv + c] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    vht[idx] = h;
    vlt[idx] = v - h;
}

__global__ void sym_finalize_kernel(const double* __restrict__ acc, long long n, int t, float* __restrict__ out, int ldo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * t) return;
    const long long row = idx / t;
    const int c = (int)(idx - row * t);
    out[row * ldo + c] = (float)acc[row * SYM_N + c];
}

size_t sym_workspace_bytes(long long n) {
    const long long npad = ((n + 127) / 128) * 128 + 64;
    return (size_t)n * SYM_N * sizeof(double) + 2 * (size_t)npad * SYM_N * sizeof(float) + 512;
}

template <int CP, int NP2>
static int run_sym(const SymArgs& a, dim3 grid, cudaStream_t st) {
    auto kernel = mvm_sym_tc_kernel<CP, NP2>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SYM_SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mvm_sym_tc_kernel)");
    kernel<<<grid, SYM_BM, SYM_SMEM_BYTES, st>>>(a);
    note_launch();
    return cuda_fail(cudaGetLastError(), "mvm_sym_tc_kernel launch");
}

int launch_sym_tc(const float* zp, long long n, int CP, const float* nlc, const float* V, int ldv, int t, float* out, int ldo,
                  int rb_begin, int rb_end, int finalize, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (workspace == nullptr || workspace_bytes < sym_workspace_bytes(n)) {
        set_error("mvm_sym: workspace %zu bytes < required %zu", workspace_bytes, sym_workspace_bytes(n));
        return ERR_WORKSPACE;
    }
    const long long npad = ((n + 127) / 128) * 128 + 64;
    char* w = (char*)workspace;
    double* acc = (double*)w;
    float* vht = (float*)(w + (((size_t)n * SYM_N * sizeof(double) + 255) / 256) * 256);
    float* vlt = vht + (size_t)npad * SYM_N;
    RPGP_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)n * SYM_N * sizeof(double), st));
    {
        const long long total = npad * SYM_N;
        sym_prep_v_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(V, n, ldv, t, npad, vht, vlt);
        note_launch();
        if (int rc = cuda_fail(cudaGetLastError(), "sym_prep_v_kernel")) return rc;
    }
    SymArgs a;
    a.z = zp; a.v = nullptr; a.vht = vht; a.vlt = vlt; a.nlc = nlc; a.acc = acc; a.n = n; a.npad = npad;
    a.nblocks = (int)((n + SYM_BM - 1) / SYM_BM);
    a.half = a.nblocks / 2 + 1;
    a.rb_begin = rb_begin; a.rb_end = rb_end;
    const int nrb = rb_end - rb_begin;
    if (nrb <= 0) return OK;
    long long want = (148LL * 2 * 16 + nrb - 1) / nrb;
    want = std::max<long long>(1, std::min<long long>(want, a.half));
    a.nsplits = (int)want;
    // the kernel reads V rows of its block from a [n][16] padded array: build it in the tail of the transposed planes' slack
    // (V is passed already padded to 16 columns by the caller when ldv == 16)
    if (ldv != SYM_N) { set_error("mvm_sym: right-hand sides must be padded to 16 columns"); return ERR_INVALID; }
    a.v = V;
    dim3 grid((unsigned)nrb, (unsigned)a.nsplits, 1);
    int rc = ERR_UNSUPPORTED;
    switch (CP) {
        case 4: rc = run_sym<4, 0>(a, grid, st); break;
        case 8: rc = run_sym<8, 1>(a, grid, st); break;
        case 12: rc = run_sym<12, 1>(a, grid, st); break;
        case 16: rc = run_sym<16, 2>(a, grid, st); break;
        case 20: rc = run_sym<20, 2>(a, grid, st); break;
        case 24: rc = run_sym<24, 2>(a, grid, st); break;
        case 28: rc = run_sym<28, 3>(a, grid, st); break;
        case 32: rc = run_sym<32, 3>(a, grid, st); break;
        default: set_error("mvm_sym: unsupported CP=%d", CP);
    }
    if (rc) return rc;
    if (finalize) {
        const long long total = n * t;
        sym_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(acc, n, t, out, ldo);
        note_launch();
        return cuda_fail(cudaGetLastError(), "sym_finalize_kernel");
    }
    return OK;
}

}  // namespace rpgp

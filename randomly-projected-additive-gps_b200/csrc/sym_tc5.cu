// sym_tc5.cu -- symmetric K(Z,Z).V with BOTH uses of every kernel value on the tensor cores.
//
// Every kernel value k(i,i') of a 128-row x 32-column tile is computed ONCE by the arithmetic warps (thread = row), split into
// its tf32 part and the remainder (3xTF32) and written to shared memory in ONE layout (SWIZZLE_128B_BASE32B, row-local stores).
// tcgen05.mma kind::tf32 then reads that single copy twice (profiles/umma_microbench_r01.txt, umma_probe3):
//     row side     out[i ,:] += sum_i' k(i,i') V[i',:]     A = S    K-major  (M = 128 rows,  K = 8 tile columns per MMA)
//     column side  out[i',:] += sum_i  k(i,i') V[i ,:]     A = S^T  MN-major (M = 64 = 32 columns of the tf32 part stacked on the
//                                                                             32 columns of the remainder, K = 8 tile rows per MMA)
// so the arithmetic warps no longer read V or spend FFMA2 issue slots on it (the kernel is issue-bound, ncu_r01_sym_cfg2_summary.md).
// The right-hand sides arrive pre-split and pre-swizzled as B operands ([16 tf32 parts | 16 remainders] x 32 rows per 4 KB block,
// written once per launch by sym5_split_rhs_kernel) and are moved by the TMA engine like the coordinate tiles.
//
// Accumulation: the row side accumulates T5_F tiles in TMEM (double-buffered by epoch), the owning arithmetic warp folds each epoch
// into Kahan-compensated FP32 registers; the column side is read back per tile by four epilogue warps (TMEM lane quadrants 0..3),
// combined through shared memory and added to FP64 accumulators with coalesced atomics.
//
// CTA = 128*HALVES arithmetic threads (setmaxnreg.inc) + 4 epilogue warps (setmaxnreg.dec).  HALVES = 1: thread = row, all 32
// columns of a tile; HALVES = 2: two threads per row, 16 columns each -> 16 arithmetic warps per SM at <= 96 registers, which
// hides the MUFU / FADD2 latencies that two warps per scheduler cannot.  Lane 0 of epilogue warps 0 and 2 also issue the row-side and
// the column-side MMAs of a tile, lane 0 of epilogue warp 1 the bulk copies (all buffers are released by the two tcgen05.commit of a tile).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "aux_kernels.cuh"
#include "dispatch.cuh"
#include "kv_kernels.cuh"
#include "sym_tc.cuh"
#include "sym_tc_dev.cuh"

namespace rpgp {

namespace {

using namespace tcdev;

constexpr int T5_F = 4;          // tiles accumulated in TMEM per row-side epoch (K = 128 per flush, like the column side)
constexpr int T5_ZST = 3;        // coordinate-tile stages
#ifndef T5_QUNROLL
#define T5_QUNROLL 1
#endif
#ifndef T5_DIAG
#define T5_DIAG 0                // diagnostic builds only (tools/sym5_variants.sh): 1 no column-side atomics, 2 no column-side MMAs, 4 no row-side MMAs
#endif
constexpr int T5_QU = T5_QUNROLL;   // unroll factor of the 4-column groups inside a tile
#ifndef T5_FENCE_ISSUER
#define T5_FENCE_ISSUER 0        // 1 (experiment): the generic -> async proxy fence for the S tile by the two issuing warps after they have
#endif                           // acquired SFULL, instead of by every arithmetic warp (MEMBAR.ALL.CTA once per tile and warp)
#ifndef T5_REGS_ARITH
#define T5_REGS_ARITH 184
#endif
#ifndef T5_REGS_HELP
#define T5_REGS_HELP 72
#endif

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t T5_S = 0;                    // S operand: buffer b at b*32768: tf32 part [128 rows][128 B], remainder 16384 B later
constexpr uint32_t T5_BC = 65536;               // B operand of the column side: V of this row block, 4 blocks x 4096 B
constexpr uint32_t T5_BT = 81920;               // B operand of the row side: V of the tile's columns, 2 stages x 4096 B
constexpr uint32_t T5_Z = 90112;                // coordinate tiles: 3 stages x 32 x CP floats (<= 4096 B each)
constexpr uint32_t T5_EPI = 102400;             // column-side epilogue exchange: 2 x [4 quadrants][16 rows][16] floats
constexpr uint32_t T5_BAR = 110592;             // mbarriers
constexpr uint32_t T5_SMEM_BYTES = T5_BAR + 256 + 1024;

// barrier indices
constexpr int B5_ZFULL = 0;      // [3]  bulk copy -> arithmetic warps
constexpr int B5_BFULL = 3;      // [2]  bulk copy -> MMA issuer (row-side B tile)
constexpr int B5_SFULL = 5;      // [2]  arithmetic warps -> MMA issuer (count 4)
constexpr int B5_TDONE = 7;      // [2]  tcgen05.commit of both issuers (count 2): S / B / z stages reusable, D2 readable, closed D1 epochs readable
constexpr int B5_EREAD = 9;      // [2]  epilogue warps have read D2 (count 4)
constexpr int B5_D1EMPTY = 11;   // [2]  arithmetic warps have folded an epoch of D1 (count 4)
constexpr int B5_BCFULL = 13;    // [1]  bulk copy of the column-side B operand


struct Sym5Args {
    const float* z;        // [nchunks][n][CP]  (blockIdx.z = coordinate chunk; the chunks' kernel values add up, so do their products)
    const float* bsplit;   // [nblocks*4][4096 B] pre-split right-hand sides (B operands)
    const float* nlc;      // [nchunks][CP or G]
    double* acc;           // [n][16] FP64 accumulators (zeroed by the launcher)
    const unsigned* gate;  // K > 1 only: bits of max |z_group|^2 written by sym_tcd.cu's pre-pass (NULL: always run)
    unsigned gate_max;     // run only when the gate is closed (tcd_gate_open false); otherwise the distance-on-tensor-core kernel
    double gate_sum4_max;  // has taken the launch
    long long n;
    int nblocks, half, nsplits, rb_begin;
};


}  // namespace

template <int CP, int KP, int G, int NP2, int HALVES, int BASE = 0>
__global__ void __launch_bounds__(128 * HALVES + 128, 2) mvm_sym_tc5_kernel(const Sym5Args a) {
    constexpr int GP = (KP == 1) ? CP : G;         // -log2c entries per chunk
    constexpr int AW = 4 * HALVES;                 // arithmetic warps
    constexpr int FC = T5_N / HALVES;              // right-hand-side columns folded by one arithmetic warp
    constexpr int REGS_ARITH = HALVES == 1 ? T5_REGS_ARITH : 96, REGS_HELP = HALVES == 1 ? T5_REGS_HELP : 48;
    if (KP > 1 && a.gate != nullptr && tcd_gate_open(a.gate, a.gate_max, a.gate_sum4_max)) return;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + T5_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + T5_BAR + 192);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* zc = a.z + (long long)blockIdx.z * a.n * CP;      // this CTA's coordinate chunk
    Tile5Iter it;
    it.I = a.rb_begin + blockIdx.x;
    it.B = a.nblocks;
    it.n = a.n;
    const int per = (a.half + a.nsplits - 1) / a.nsplits;
    it.k_begin = blockIdx.y * per;
    it.ntiles = 4 * (min(a.half, it.k_begin + per) - it.k_begin);
    if (it.ntiles < 0) it.ntiles = 0;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < T5_ZST; ++s) mbar_init(&bars[B5_ZFULL + s], 1);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars[B5_BFULL + b], 1);
            mbar_init(&bars[B5_SFULL + b], AW);
            mbar_init(&bars[B5_TDONE + b], 2);
            mbar_init(&bars[B5_EREAD + b], 4);
            mbar_init(&bars[B5_D1EMPTY + b], AW);
        }
        mbar_init(&bars[B5_BCFULL], 1);
        mbar_fence_init();
    }
    if (warp == AW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;   // D1[epoch & 1] (row side, 128 lanes x 32 columns) at column 32*(epoch&1); D2[b] at 64 + 32*b

    if (warp < AW) {
        // =========================================== arithmetic warps ===================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_ARITH));
        const int rtid = tid & (T5_ROWS - 1), half = tid >> 7;     // row in the block; which part of the tile's columns
        const long long row = (long long)it.I * T5_ROWS + rtid;
        const bool valid = row < a.n;
        RowCoords<CP, KP, G> r;
        load_row_coords<CP, KP, G, BASE>(r, zc + row * CP, valid, a.nlc + (int)blockIdx.z * GP);
        f32x2 acc[FC / 2], comp[FC / 2];
#pragma unroll
        for (int q = 0; q < FC / 2; ++q) { acc[q] = 0ull; comp[q] = 0ull; }

        // fold one closed epoch of the row-side accumulator (TMEM lanes of this warp's quadrant = its rows; with two threads per
        // row each takes half of the right-hand sides) into the running total
        auto fold_epoch = [&](int e) {
            float d[FC], x[FC];
            const uint32_t ta = tmem + 32u * (uint32_t)(e & 1) + (uint32_t)(half * FC) + ((uint32_t)((warp & 3) * 32) << 16);
            if constexpr (FC == 16) { tmem5_ld16(ta, d); tmem5_ld16(ta + 16u, x); }    // Sh.Vh + Sl.Vh | Sh.Vl
            else { tmem5_ld8(ta, d); tmem5_ld8(ta + 16u, x); }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar5_arrive(&bars[B5_D1EMPTY + (e & 1)]);
#pragma unroll
            for (int q = 0; q < FC / 2; ++q) {
                const f32x2 y = sub2(pack2(d[2 * q] + x[2 * q], d[2 * q + 1] + x[2 * q + 1]), comp[q]);
                const f32x2 tsum = add2(acc[q], y);
                comp[q] = sub2(sub2(tsum, acc[q]), y);
                acc[q] = tsum;
            }
        };

        int j = 0, folded = 0;
        for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++j) {
            const int zs = j % T5_ZST, b = j & 1;
            const long long c0 = it.col0(t);
            const int cols = (int)min((long long)T5_BN, a.n - c0);
            mbar_wait(&bars[B5_ZFULL + zs], (uint32_t)((j / T5_ZST) & 1));
            if (j >= 2) {   // tile j-2 has left the tensor core: S buffer b is free, and every epoch that ended at or before j-2 is closed
                mbar_wait(&bars[B5_TDONE + b], (uint32_t)(((j >> 1) - 1) & 1));
                tc5_fence_after();
                if (j >= T5_F + 1 && (j - 1) % T5_F == 0) fold_epoch(folded++);
            }
            const float* zt = reinterpret_cast<const float*>(sm + T5_Z) + (size_t)zs * T5_BN * CP;
            unsigned char* sc = sm + T5_S + (uint32_t)b * 32768u;
            auto tile_body = [&](auto full_tile) {
                constexpr bool FULL = decltype(full_tile)::value;
#pragma unroll T5_QU
                for (int q = half * (T5_BN / 4 / HALVES); q < (half + 1) * (T5_BN / 4 / HALVES); ++q) {
                    float sv[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 4 * q + e;
                        const float val = pair_kernel_value<CP, KP, G, NP2, BASE>(r, zt + c * CP);
                        sv[e] = (FULL || c < cols) ? val : 0.f;      // columns behind a partial tile hold stale bytes
                    }
                    // split, SWIZZLE_128B_BASE32B (32-byte chunks ^ row % 4), row-local stores
                    float4 h, l;
                    h.x = tf32_hi5(sv[0]); h.y = tf32_hi5(sv[1]); h.z = tf32_hi5(sv[2]); h.w = tf32_hi5(sv[3]);
                    l.x = sv[0] - h.x; l.y = sv[1] - h.y; l.z = sv[2] - h.z; l.w = sv[3] - h.w;
                    const uint32_t off = (uint32_t)rtid * 128u + (((((uint32_t)q >> 1) ^ ((uint32_t)rtid & 3u)) << 5) | (((uint32_t)q & 1u) << 4));
                    *reinterpret_cast<float4*>(sc + off) = h;
                    *reinterpret_cast<float4*>(sc + 16384u + off) = l;
                }
            };
            if (cols == T5_BN) tile_body(std::true_type{}); else tile_body(std::false_type{});
            if (!T5_FENCE_ISSUER) fence5_async_smem();
            __syncwarp();
            if (lane == 0) mbar5_arrive(&bars[B5_SFULL + b]);
        }
        if (j > 0) {
            mbar_wait(&bars[B5_TDONE + ((j - 1) & 1)], (uint32_t)(((j - 1) >> 1) & 1));   // the last commit covers every earlier MMA
            tc5_fence_after();
            const int epochs = (j + T5_F - 1) / T5_F;
            while (folded < epochs) fold_epoch(folded++);
            if (valid) {
                double* dst = a.acc + row * T5_N + half * FC;
#pragma unroll
                for (int q = 0; q < FC / 2; ++q) {
                    float x, y;
                    unpack2(acc[q], x, y);
                    atomicAdd(dst + 2 * q, (double)x);
                    atomicAdd(dst + 2 * q + 1, (double)y);
                }
            }
        }
    } else {
        // =========================================== epilogue warps (+ MMA issue, + bulk copies) ========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_HELP));
        const int qd = warp - AW;                                  // TMEM lane quadrant (AW is a multiple of 4)
        // row-side MMAs from quadrant 0's warp, column-side MMAs from quadrant 2's, bulk copies from quadrant 1's.  Every lane of the
        // warp follows the barriers and keeps the counters; the tcgen05.mma / commit / cp.async.bulk themselves are issued by one
        // elected lane (elect_one, sym_tc_dev.cuh): warp-uniform control flow keeps the descriptors in uniform registers and the
        // UTCHMMAs back to back (under `lane == 0` ptxas wraps each of them in a waterfall loop, ~150 clk per MMA)
        const bool row_issuer = (qd == 0), col_issuer = (qd == 2), loader = (qd == 1);
        constexpr uint32_t IDESC_ROW_N32 = idesc5_tf32(128, 2 * T5_N, 0, 0), IDESC_ROW_N16 = idesc5_tf32(128, T5_N, 0, 0);
        constexpr uint32_t IDESC_COL = idesc5_tf32(64, 2 * T5_N, 1, 0);
        const unsigned char* bsplit = reinterpret_cast<const unsigned char*>(a.bsplit);

        // the loader runs ahead of the consumers: coordinate tiles three deep, B tiles two deep
        int tz = it.next_live(0), jz = 0, tb = tz, jb = 0;
        auto load_z = [&]() {
            const long long c0 = it.col0(tz);
            const uint32_t cols = (uint32_t)min((long long)T5_BN, a.n - c0);
            const int zs = jz % T5_ZST;
            if (elect_one()) {
                mbar_expect_tx(&bars[B5_ZFULL + zs], cols * CP * (uint32_t)sizeof(float));
                bulk_g2s(sm + T5_Z + (uint32_t)zs * T5_BN * CP * 4u, zc + c0 * CP, cols * CP * (uint32_t)sizeof(float), &bars[B5_ZFULL + zs]);
            }
            __syncwarp();
            tz = it.next_live(tz + 1);
            ++jz;
        };
        auto load_b = [&]() {
            const long long c0 = it.col0(tb);
            const int bs = jb & 1;
            if (elect_one()) {
                mbar_expect_tx(&bars[B5_BFULL + bs], 4096u);
                bulk_g2s(sm + T5_BT + (uint32_t)bs * 4096u, bsplit + (c0 / T5_BN) * 4096, 4096u, &bars[B5_BFULL + bs]);
            }
            __syncwarp();
            tb = it.next_live(tb + 1);
            ++jb;
        };
        if (loader && tz < it.ntiles) {   // (a CTA without tiles must not leave copies in flight)
            if (elect_one()) {
                mbar_expect_tx(&bars[B5_BCFULL], 16384u);
                bulk_g2s(sm + T5_BC, bsplit + (long long)it.I * 16384, 16384u, &bars[B5_BCFULL]);
            }
            __syncwarp();
            for (int s = 0; s < T5_ZST && tz < it.ntiles; ++s) load_z();
            for (int s = 0; s < 2 && tb < it.ntiles; ++s) load_b();
        }

        int j = 0, jc = 0;
        for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++j) {
            const int b = j & 1;
            const bool diag = it.diag(t);
            if (row_issuer) {   // D1 += Sh.[Vh|Vl] + Sl.Vh, four k-steps of 8 tile columns
                const int e = j / T5_F;
                mbar5_wait_sleep(&bars[B5_BFULL + b], (uint32_t)((j >> 1) & 1));
                if (j % T5_F == 0 && e >= 2) mbar5_wait_sleep(&bars[B5_D1EMPTY + (e & 1)], (uint32_t)(((e >> 1) - 1) & 1));
                mbar5_wait_sleep(&bars[B5_SFULL + b], (uint32_t)((j >> 1) & 1));
                if (T5_FENCE_ISSUER) fence5_async_smem();
                tc5_fence_after();
                if (elect_one()) {
                    if (!(T5_DIAG & 4)) {
                        const uint32_t sbuf = base + T5_S + (uint32_t)b * 32768u;
                        const uint32_t d1 = tmem + 32u * (uint32_t)(e & 1);
                        const uint64_t dA_h = smem_desc5(sbuf, 512, 512, LAYOUT5_SW128_BASE32B);
                        const uint64_t dA_l = smem_desc5(sbuf + 16384u, 512, 512, LAYOUT5_SW128_BASE32B);
                        const uint64_t dB = smem_desc5(base + T5_BT + (uint32_t)b * 4096u, 16, 1024, LAYOUT5_SW128);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            umma5(d1, dA_h + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC_ROW_N32, (j % T5_F != 0 || ks > 0) ? 1u : 0u);
                            umma5(d1, dA_l + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC_ROW_N16, 1u);
                        }
                    }
                    umma5_commit(&bars[B5_TDONE + b]);
                }
                __syncwarp();
            }
            if (col_issuer) {   // D2 = [Sh ; Sl]^T . [Vh|Vl], sixteen k-steps of 8 tile rows (nothing to do on the diagonal block)
                if (j == 0) mbar5_wait_sleep(&bars[B5_BCFULL], 0u);
                if (j >= 2) mbar5_wait_sleep(&bars[B5_EREAD + b], (uint32_t)(((j >> 1) - 1) & 1));        // D2[b] has been read
                mbar5_wait_sleep(&bars[B5_SFULL + b], (uint32_t)((j >> 1) & 1));
                if (T5_FENCE_ISSUER) fence5_async_smem();
                tc5_fence_after();
                if (elect_one()) {
                    if (!diag && !(T5_DIAG & 2)) {
                        const uint32_t d2 = tmem + 64u + 32u * (uint32_t)b;
                        const uint64_t dA = smem_desc5(base + T5_S + (uint32_t)b * 32768u, 16384, 512, LAYOUT5_SW128_BASE32B);
                        const uint64_t dB = smem_desc5(base + T5_BC, 16, 1024, LAYOUT5_SW128);
#pragma unroll
                        for (int g = 0; g < 16; ++g)
                            umma5(d2, dA + (uint64_t)((g * 1024) >> 4), dB + (uint64_t)(((g >> 2) * 4096 + (g & 3) * 32) >> 4), IDESC_COL, g > 0 ? 1u : 0u);
                    }
                    umma5_commit(&bars[B5_TDONE + b]);
                }
                __syncwarp();
            }
            __syncwarp();
            mbar5_wait_sleep(&bars[B5_TDONE + b], (uint32_t)((j >> 1) & 1));
            tc5_fence_after();
            if (loader) {   // tile j is through: its coordinate stage, S buffer and B stage are free
                if (tz < it.ntiles) load_z();
                if (tb < it.ntiles) load_b();
            }
            // quadrants 0,1 hold the tf32-part rows of columns 0..15 / 16..31 (lanes 0..15), quadrants 2,3 the remainder rows
            float* P = reinterpret_cast<float*>(sm + T5_EPI) + (jc & 1) * 1024;
            if (!diag) {
                const uint32_t ta = tmem + 64u + 32u * (uint32_t)b + ((uint32_t)(qd * 32) << 16);
#pragma unroll
                for (int ch = 0; ch < 2; ++ch) {   // 8 right-hand sides at a time: [Vh part | Vl part] summed
                    float d[8], x[8];
                    tmem5_ld8(ta + 8u * ch, d);
                    tmem5_ld8(ta + 16u + 8u * ch, x);
                    if (lane < 16) {
                        float4* dst = reinterpret_cast<float4*>(P + qd * 256 + lane * 16 + 8 * ch);
                        dst[0] = make_float4(d[0] + x[0], d[1] + x[1], d[2] + x[2], d[3] + x[3]);
                        dst[1] = make_float4(d[4] + x[4], d[5] + x[5], d[6] + x[6], d[7] + x[7]);
                    }
                }
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar5_arrive(&bars[B5_EREAD + b]);
            if (!diag) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int et = tid - 32 * AW, c = et & 15;
                const long long c0 = it.col0(t);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rr = (et >> 4) + 8 * k;      // tile column 0..31 = output row c0 + rr
                    const float v = P[(rr >> 4) * 256 + (rr & 15) * 16 + c] + P[(2 + (rr >> 4)) * 256 + (rr & 15) * 16 + c];
                    if (c0 + rr < a.n && !(T5_DIAG & 1)) atomicAdd(a.acc + (c0 + rr) * T5_N + c, (double)v);
                }
                ++jc;
            }
        }
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == AW) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
    }
}

template <int CP, int KP, int G, int NP2, int HALVES, int BASE = 0>
static int run_sym5(const Sym5Args& a, dim3 grid, cudaStream_t st) {
    auto kernel = mvm_sym_tc5_kernel<CP, KP, G, NP2, HALVES, BASE>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T5_SMEM_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mvm_sym_tc5_kernel)");
    kernel<<<grid, 128 * HALVES + 128, T5_SMEM_BYTES, st>>>(a);
    note_launch();
    return cuda_fail(cudaGetLastError(), "mvm_sym_tc5_kernel launch");
}

// V16 [n][16] -> B operands: per 32 rows one 4096-byte block [16 tf32 parts | 16 remainders][32 rows], K-major SWIZZLE_128B;
// rows >= n are zero (n_pad = 128 * nblocks rows)
__global__ void sym5_split_rhs_kernel(const float* __restrict__ v16, long long n, long long n_pad, float* __restrict__ bsplit) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pad * T5_N) return;
    const int c = (int)(idx / n_pad);               // consecutive threads = consecutive rows: 128-byte runs in the output
    const long long row = idx - (long long)c * n_pad;
    const float v = row < n ? __ldg(v16 + row * T5_N + c) : 0.f;
    const float h = tf32_hi5(v);
    unsigned char* blk = reinterpret_cast<unsigned char*>(bsplit) + (row >> 5) * 4096;
    const uint32_t off = (uint32_t)(c >> 3) * 1024u + sw128_5((uint32_t)(c & 7), (uint32_t)(row & 31));
    *reinterpret_cast<float*>(blk + off) = h;
    *reinterpret_cast<float*>(blk + 2048u + off) = v - h;
}

__global__ void sym5_finalize_kernel(const double* __restrict__ acc, long long n, int t, float* __restrict__ out, int ldo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * t) return;
    const long long row = idx / t;
    const int c = (int)(idx - row * t);
    out[row * ldo + c] = (float)acc[row * T5_N + c];
}

int launch_sym_tc5(const float* zp, long long n, const Layout& lay, const float* nlc, const float* V16, int t, float* out, int ldo,
                   int rb_begin, int rb_end, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const int CP = lay.CP, KP = lay.KP, G = lay.G;
    const int nblocks = (int)((n + T5_ROWS - 1) / T5_ROWS);
    const size_t acc_bytes = ((size_t)n * T5_N * sizeof(double) + 1023) & ~(size_t)1023;
    const size_t base_bytes = sym_base_workspace_bytes(n);
    const size_t need = acc_bytes + (size_t)nblocks * 16384;
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("mvm_sym: workspace %zu bytes < required %zu", workspace_bytes, need);
        return ERR_WORKSPACE;
    }
    double* acc = (double*)workspace;
    float* bsplit = (float*)((unsigned char*)workspace + acc_bytes);
    RPGP_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)n * T5_N * sizeof(double), st));
    const int nrb = rb_end - rb_begin;
    if (nrb > 0) {
        const long long n_pad = (long long)nblocks * T5_ROWS;
        sym5_split_rhs_kernel<<<(unsigned)((n_pad * T5_N + 255) / 256), 256, 0, st>>>(V16, n, n_pad, bsplit);
        note_launch();
        RPGP_CUDA_OK(cudaGetLastError());
        Sym5Args a;
        a.z = zp; a.bsplit = bsplit; a.nlc = nlc; a.acc = acc; a.n = n;
        a.gate = nullptr; a.gate_max = 0; a.gate_sum4_max = 0.0;
        a.nblocks = nblocks;
        a.half = nblocks / 2 + 1;
        a.rb_begin = rb_begin;
        static const int splits_env = [] { const char* e = getenv("RPGP_SYM_SPLITS"); return e ? atoi(e) : 0; }();
        long long want = pick_splits((long long)nrb * lay.nchunks, a.half, 148 * 2);      // two CTAs per SM
        if (splits_env > 0) want = splits_env;
        want = std::max<long long>(1, std::min<long long>(want, a.half));
        a.nsplits = (int)want;
        dim3 grid((unsigned)nrb, (unsigned)a.nsplits, (unsigned)lay.nchunks);
        int rc = ERR_UNSUPPORTED;
        if (KP == 1 && lay.base != 0) {
            // K = 1 with the Matern-1.5 / inverse-multiquadric / cosine base kernel: one MUFU per projection (no square root: the distance is |d|)
#define RPGP_SYM5_K1B(CPv)                                                                                             \
            if (CP == CPv) {                                                                                           \
                if constexpr (CPv <= 24)                                                                               \
                    rc = lay.base == BASE_MATERN15 ? run_sym5<CPv, 1, CPv, 0, 2, BASE_MATERN15>(a, grid, st)           \
                         : lay.base == BASE_IMQ    ? run_sym5<CPv, 1, CPv, 0, 2, BASE_IMQ>(a, grid, st)                \
                                                   : run_sym5<CPv, 1, CPv, 0, 2, BASE_COS>(a, grid, st);               \
                else                                                                                                   \
                    rc = lay.base == BASE_MATERN15 ? run_sym5<CPv, 1, CPv, 0, 1, BASE_MATERN15>(a, grid, st)           \
                         : lay.base == BASE_IMQ    ? run_sym5<CPv, 1, CPv, 0, 1, BASE_IMQ>(a, grid, st)                \
                                                   : run_sym5<CPv, 1, CPv, 0, 1, BASE_COS>(a, grid, st);               \
            }
            RPGP_SYM5_K1B(4) RPGP_SYM5_K1B(8) RPGP_SYM5_K1B(12) RPGP_SYM5_K1B(16) RPGP_SYM5_K1B(20) RPGP_SYM5_K1B(24) RPGP_SYM5_K1B(28) RPGP_SYM5_K1B(32)
#undef RPGP_SYM5_K1B
            if (rc == ERR_UNSUPPORTED) set_error("mvm_sym: no K=1 kernel for CP=%d base=%d", CP, lay.base);
        } else if (KP == 1) {
            // polynomial-exp2 pairs (kv_kernels.cuh::exp2_neg_poly2); RPGP_SYM_POLY_PAIRS overrides (tools sweep)
            static const int np_env = [] { const char* e = getenv("RPGP_SYM_POLY_PAIRS"); return e ? atoi(e) : -1; }();
            const int np = np_env >= 0 ? np_env : (CP >= 20 ? 2 : (CP >= 16 ? 1 : 0));
            // two threads per row where the row's coordinates leave room in 96 registers; RPGP_SYM_HALVES overrides (tools sweep)
            static const int halves_env = [] { const char* e = getenv("RPGP_SYM_HALVES"); return e ? atoi(e) : 0; }();
            const int halves = halves_env ? halves_env : (CP <= 24 ? 2 : 1);
#define RPGP_SYM5_CASE(CPv, NPv)                                                                                       \
            if (CP == CPv && np == NPv) {                                                                              \
                if constexpr (CPv <= 24) rc = halves == 2 ? run_sym5<CPv, 1, CPv, NPv, 2>(a, grid, st) : run_sym5<CPv, 1, CPv, NPv, 1>(a, grid, st); \
                else rc = run_sym5<CPv, 1, CPv, NPv, 1>(a, grid, st);                                                  \
            }
            RPGP_SYM5_CASE(4, 0) RPGP_SYM5_CASE(8, 0) RPGP_SYM5_CASE(12, 0) RPGP_SYM5_CASE(16, 0) RPGP_SYM5_CASE(16, 1)
            RPGP_SYM5_CASE(20, 0) RPGP_SYM5_CASE(20, 1) RPGP_SYM5_CASE(20, 2) RPGP_SYM5_CASE(20, 3) RPGP_SYM5_CASE(24, 0) RPGP_SYM5_CASE(24, 2)
            RPGP_SYM5_CASE(28, 0) RPGP_SYM5_CASE(28, 2) RPGP_SYM5_CASE(32, 0) RPGP_SYM5_CASE(32, 2)
#undef RPGP_SYM5_CASE
            if (rc == ERR_UNSUPPORTED) set_error("mvm_sym: no kernel for CP=%d poly pairs=%d (compiled: 0 for every CP, 1 for CP 16/20, 2 for CP >= 20)", CP, np);
        } else {
            // K > 1, squared distances on the tensor cores (sym_tcd.cu) while the coordinates are small enough for the cancellation
            // in |z|^2 + |z'|^2 - 2 z.z'; the direct-difference kernel below then returns at once (and the other way round)
            if (lay.base == 0 && plan_tcd(lay).supported && workspace_bytes >= base_bytes + tcd_workspace_bytes(n, lay)) {
                const unsigned* gate = nullptr;
                if (int rcd = launch_sym_tcd(zp, n, lay, nlc, bsplit, acc, nblocks, rb_begin, nrb, (unsigned char*)workspace + base_bytes,
                                             workspace_bytes - base_bytes, &gate, st))
                    return rcd;
                a.gate = gate;
                const TcdGate gb = tcd_gate(n, lay);
                a.gate_max = gb.max_bits;
                a.gate_sum4_max = gb.sum4_max;
            }
            // K > 1: the (KP, G, CP) chunk shapes of dispatch.cuh; one MUFU per group, so no polynomial offload; two threads per row
            // (Matern-1.5 and the inverse multiquadric: the same kernel with another function of the group's squared distance)
#define RPGP_SYM5_KN(KPv, Gv, CPv, TPv)                                                                                \
            if (KP == KPv && G == Gv && CP == CPv)                                                                     \
                rc = lay.base == BASE_MATERN15 ? run_sym5<CPv, KPv, Gv, 0, 2, BASE_MATERN15>(a, grid, st)              \
                     : lay.base == BASE_IMQ    ? run_sym5<CPv, KPv, Gv, 0, 2, BASE_IMQ>(a, grid, st)                   \
                                               : run_sym5<CPv, KPv, Gv, 0, 2, 0>(a, grid, st);
            RPGP_KN_SHAPE_LIST(RPGP_SYM5_KN, 0)
#undef RPGP_SYM5_KN
            if (rc == ERR_UNSUPPORTED) set_error("mvm_sym: no kernel for chunk shape KP=%d G=%d CP=%d", KP, G, CP);
        }
        if (rc) return rc;
    }
    const long long total = n * t;
    sym5_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(acc, n, t, out, ldo);
    note_launch();
    return cuda_fail(cudaGetLastError(), "sym5_finalize_kernel");
}

}  // namespace rpgp

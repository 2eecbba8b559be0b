// sym_tcd.cu -- symmetric K(Z,Z).V for K > 1 coordinates per projection group with the squared distances on the tensor cores.
//
// For a group of K coordinates the exponent of the kernel value is
//     U[i,i'] = -log2 c + |z_i|^2 + |z_i'|^2 - 2 z_i.z_i'          (coordinates pre-scaled so that k = 2^-U)
// and sym_tc5.cu spends 2K+1 FP32 lane-operations per (pair, group) on it -- for K = 5 / 20 the FMA pipe and the broadcast loads
// of the column coordinates bind long before the XU pipe does.  Here U is an inner product of augmented vectors, six coordinates
// per k-step of eight:
//     A_i  = [ -2 z_i[6s..6s+5] , |z_i[6s..]|^2 (- log2 c in step 0) , 1 ]        B_i' = [ z_i'[6s..6s+5] , 1 , |z_i'[6s..]|^2 ]
// evaluated by tcgen05.mma kind::tf32 with the 3xTF32 split (Ah.Bh + Al.Bh + Ah.Bl, FP32 accumulation in TMEM): the arithmetic
// warps only read U back (tcgen05.ld, lane = row), take the exponential and add up the groups.  Everything downstream of the
// kernel value -- the split of S, S.V on the row side, S^T.V on the column side, the FP64 accumulators -- is sym_tc5.cu's scheme.
//
// Accuracy.  The cancellation in U costs an absolute error proportional to |z|^2.  What keeps it small (tools/tcd_check.py adv,
// profiles/tcd_accuracy_r01.txt): coordinates are centred on their column means (differences do not change); the split is
// round-to-nearest for both parts (a truncating split biases every U by ~3e-7 |z|^2); every k-step carries the partial norms of
// ITS six coordinates, so that the running sum in TMEM is a partial squared distance -- small for exactly the pairs that matter;
// pairs i == i' get their exact value U = -log2 c.  The pre-pass records max |z - mean|^2 over rows and groups and the kernel runs
// only while that stays below a bound (rpgp::tcd_gate_bound); otherwise the direct-difference kernel of sym_tc5.cu takes the launch
// -- both are launched, one of them returns at once.
//
// Operands: the pre-pass (tcd_build_images_kernel) writes, per 128-row block and per 32-column tile, the exact shared-memory bytes
// of the A / B operands (K-major, SWIZZLE_128B, tf32 part and remainder planes; a 128-byte line holds 4 k-steps) so that one bulk
// copy per tile delivers a ready operand.  A chunk holds GT groups of KS = ceil(K/6) k-steps (GT*KS <= 4*NL, NL <= 2).
//
// CTA (one per SM, 24 warps):
//   * 16 arithmetic warps in two teams of eight that take the tiles in turn (team = tile parity = S buffer).  Measured (round 2,
//     profiles/tcd_timeline_r02.txt): the teams run IN STEP -- they share the XU pipe during the exponentials and are in their
//     hand-off code at the same time -- so that code is kept as short as it can be (precomputed toggling addresses, no S2R / LDC in
//     the loops, the diagonal fix and the odd last batch in their own loop copies);
//   * 4 epilogue warps (TMEM lane quadrants) for the column side;
//   * 4 helper warps, one elected lane each (warp-uniform control flow: sym_tc_dev.cuh elect_one): the S-side MMAs (S.V then S^T.V of
//     a tile), the bulk copies, and the distance MMAs of team 0's / team 1's tiles.  One distance issuer PER TEAM (round 2): with a
//     single in-order stream a team whose buffers are full stalls the other team's exponents as well; the tensor pipe serialises the
//     MMAs either way (40 clk per M128 N32 K8 MMA whatever the number of issuing warps, profiles/umma_rate_r02.txt).  The copy warp
//     refills a B-image stage as soon as the tile's distance MMAs have completed (ZFREE, committed by the distance issuer), not when
//     the whole tile is through.
// TMEM: D1 (row side) 2 x 32 columns, D2 (column side) 2 x 32, the A operand of the distance MMAs 128, D0 (exponents) 2 teams x 2
// buffers x 2 groups x 32 = all 512 columns.  Registers: setmaxnreg gives the arithmetic warps 96, the epilogue and helper warps 48.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "aux_kernels.cuh"
#include "sym_tc.cuh"
#include "sym_tc_dev.cuh"

namespace rpgp {

namespace {

using namespace tcdev;

constexpr int D_F = 4;            // tiles accumulated in TMEM per row-side epoch
#ifndef TCD_ZST
#define TCD_ZST 3
#endif
constexpr int D_ZST = TCD_ZST;    // B-image stages.  With an EVEN number a stage only ever holds tiles of one team, and a distance issuer no
                                  // longer has to walk through the other team's B-image barrier phases -- the one wait of this kernel
                                  // that can alias under unbounded starvation of a helper warp (tests/test_tcd_protocol_cpu.py).
                                  // -DTCD_ZST=4 costs NL x 8 KB of shared memory; not yet measured on a GPU, hence not the default.
constexpr int D_AW = 16;          // arithmetic warps
constexpr int D_FC = 4;            // right-hand-side columns folded per arithmetic thread (16 warps: 4 lane quadrants x 4 column parts)
constexpr int D_NDI = 2;          // distance-MMA issuing warps (one per team)
#ifndef TCD_NBUF
#define TCD_NBUF 2
#endif
constexpr int D_NBUF = TCD_NBUF;  // D0 buffers (batches of two groups) per team
constexpr int D_ISSUERS = 2 + D_NDI; // row side, column side, distances
constexpr int D_THREADS = 32 * (D_AW + 4 + D_ISSUERS);
#ifndef TCD_DIAG
#define TCD_DIAG 0                 // diagnostic builds only (tools/tcd_diag.sh): 1 no column atomics, 2 no column MMAs, 4 no row MMAs,
#endif                             // 8 no distance MMAs, 32 no D2 reads; -DTCD_DEBUG_STAMPS adds clock64 stamps of one CTA (RPGP_TCD_DBG=1 prints them)
#ifndef TCD_SLEEP_NS
#define TCD_SLEEP_NS 64
#endif
constexpr int D_SLEEP = TCD_SLEEP_NS;   // back-off of the helper warps' barrier polls (ns)
constexpr size_t D_WS_HEADER = 16384;   // workspace header: gate word, then column sums (doubles) from byte 1024
constexpr float D_PAD = 16384.f;  // exponent of padding rows / groups: 2^-16384 == 0

// shared-memory map (bytes from a 1024-aligned base)
constexpr uint32_t D_S = 0;                    // S operand: buffer b at b*32768: tf32 part [128 rows][128 B], remainder 16384 B later
constexpr uint32_t D_BC = 65536;               // B operand of the column side: V of this row block, 4 blocks x 4096 B
constexpr int D_BST = 4;                       // stages of the row-side B operand
constexpr uint32_t D_BT = 81920;               // B operand of the row side: V of the tile's columns, D_BST stages x 4096 B
constexpr uint32_t D_EPI = 98304;              // column-side epilogue exchange: 2 x [4 quadrants][16 rows][16] floats
constexpr uint32_t D_BIMG = 106496;            // B operand of the distance MMAs: D_ZST stages x NL x [hi|lo] x [32][128 B]
__host__ __device__ constexpr uint32_t d_bimg(int NL) { return D_BIMG; }
__host__ __device__ constexpr uint32_t d_bar(int NL) { return d_bimg(NL) + (uint32_t)D_ZST * NL * 8192u; }
__host__ __device__ constexpr uint32_t d_smem_bytes(int NL) { return d_bar(NL) + 512 + 1024; }

// TMEM columns
// TMEM columns: D1 (row side) 2 x 32, D2 (column side) 2 x 32, the A operand of the distance MMAs (this row block's augmented
// coordinates, written once per CTA by tcgen05.st: k-step s at 8 s, tf32 parts then remainders) 2 x 4 NL x 8 <= 128, D0 2 teams x D_NBUF x 64
constexpr uint32_t D_TM_D1 = 0, D_TM_D2 = 64, D_TM_A = 128, D_TM_D0 = 256;
static_assert(D_TM_D0 + 2 * D_NBUF * 64 <= 512, "TMEM columns");

// barrier indices (32 x 8 bytes)
constexpr int BD_ZFULL = 0;                    // [D_ZST]  bulk copy of a B image -> distance issuers
constexpr int BD_ZFREE = BD_ZFULL + D_ZST;     // [D_ZST]  tcgen05.commit of the tile's distance issuer: the B image has been consumed
constexpr int BD_BFULL = BD_ZFREE + D_ZST;     // [D_BST]  bulk copy -> S-side issuer (B tile of V)
constexpr int BD_SFULL = BD_BFULL + D_BST;     // [2]  the tile's team -> S-side issuer (count 8)
constexpr int BD_TDONE = BD_SFULL + 2;         // [2]  tcgen05.commit of the S-side issuer after a tile's row-side and column-side MMAs
constexpr int BD_EREAD = BD_TDONE + 2;         // [2]  epilogue warps have read D2 (count 4)
constexpr int BD_D1EMPTY = BD_EREAD + 2;       // [2]  arithmetic warps have folded an epoch of D1 (count 16)
constexpr int BD_BCFULL = BD_D1EMPTY + 2;      // [1]  bulk copy of the column-side B operand
constexpr int BD_AFULL = BD_BCFULL + 1;        // [1]  the A image is in tensor memory (count 4: the warps of rows 0..127)
constexpr int BD_D0FULL = BD_AFULL + 1;        // [2 teams][D_NBUF]  tcgen05.commit of the team's distance issuer for a batch of two groups
constexpr int BD_D0FREE = BD_D0FULL + 2 * D_NBUF;   // [2 teams][D_NBUF]  the team's warps have read the batch (count 8)
static_assert(BD_D0FREE + 2 * D_NBUF <= 32, "barrier block is 256 bytes");
#ifndef TCD_FENCE_ISSUER
#define TCD_FENCE_ISSUER 0       // 0: every arithmetic warp fences its S stores (generic -> async proxy) before SFULL; 1 (experiment,
#endif                           // same speed, parity tests pass): one fence by the S-side issuer after it has acquired SFULL

__device__ __forceinline__ void tmem5_ld4(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = __uint_as_float(r[q]);
}

// four loads in flight, one wait
__device__ __forceinline__ void tmem5_ld8x4(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, float* v0, float* v1, float* v2, float* v3) {
    uint32_t r[32];
    const uint32_t ta[4] = {t0, t1, t2, t3};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[8 * k]), "=r"(r[8 * k + 1]), "=r"(r[8 * k + 2]), "=r"(r[8 * k + 3]), "=r"(r[8 * k + 4]), "=r"(r[8 * k + 5]),
                       "=r"(r[8 * k + 6]), "=r"(r[8 * k + 7])
                     : "r"(ta[k]));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        v0[q] = __uint_as_float(r[q]); v1[q] = __uint_as_float(r[8 + q]); v2[q] = __uint_as_float(r[16 + q]); v3[q] = __uint_as_float(r[24 + q]);
    }
}

__device__ __forceinline__ void mbar_wait_ns(int ns, uint64_t* bar, uint32_t parity) {
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
        if (ns > 0) __nanosleep(ns);
    }
}

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

struct SymDArgs {
    const unsigned char* aimg;   // [nchunks][nblocks][NL][2][128][128 B]
    const unsigned char* bimg;   // [nchunks][nblocks*4][NL][2][32][128 B]
    const float* bsplit;         // [nblocks*4][4096 B] pre-split right-hand sides
    const float* nlc;            // [J] -log2 c per group
    double* acc;                 // [n][16] FP64 accumulators
    const unsigned* gate;        // pre-pass statistics of the centred squared group norms (sym_tc_dev.cuh: tcd_gate_open)
    unsigned gate_max;
    double gate_sum4_max;
    long long n;
    int nblocks, half, nsplits, rb_begin;
    int G, KS, NB, J;            // groups per chunk, k-steps per group, D0 batches (of two groups) per tile, total groups
    unsigned* dbg;               // RPGP_TCD_DBG (TCD_DEBUG_STAMPS builds): clock stamps of one CTA, [DBG_NT tiles][DBG_NC]
    int sleep_ns;                // back-off of the helper warps' barrier polls
};

}  // namespace

// -DTCD_DEBUG_STAMPS: one CTA keeps clock stamps of DBG_NT tiles in shared memory (a store to global memory would sit in front of the
// arithmetic warps' proxy fence, which waits for it) and copies them out at the end; RPGP_TCD_DBG=1 prints them (launch_sym_tcd).
// columns: 0 top | 1-4 D0FULL passed, batch k | 5-8 exponentials of batch k issued | 9 TDONE passed | 10 folded | 11 S stores issued |
// 12 proxy fence done | 13 SFULL arrive || distance issuer: 14-17 batch k issued, 18 ZFULL passed || S-side issuer: 19 waits passed,
// 20 committed || epilogue: 21 TDONE seen, 22 done
#ifdef TCD_DEBUG_STAMPS
constexpr int DBG_T0 = 16, DBG_NT = 40, DBG_NC = 24;
#define TCD_STAMP(j, k) do { if (stamping && (j) >= DBG_T0 && (j) < DBG_T0 + DBG_NT) dbg_sm[((j) - DBG_T0) * DBG_NC + (k)] = (uint32_t)clock64(); } while (0)
#else
#define TCD_STAMP(j, k) do { } while (0)
#endif
#ifndef TCD_REGS_ARITH
#define TCD_REGS_ARITH 96
#endif
#ifndef TCD_REGS_EPI
#define TCD_REGS_EPI 48
#endif
#ifndef TCD_REGS_HELP
#define TCD_REGS_HELP 48
#endif
// setmaxnreg moves registers inside the CTA's OWN allocation (768 threads x D_REGS_LAUNCH): a split that needs more than that spins
// in USETMAXREG.TRY_ALLOC forever (measured the hard way: 104 / 56 / 40 from a launch at 72 hangs)
constexpr int D_REGS_LAUNCH = 80, D_REGS_ARITH = TCD_REGS_ARITH, D_REGS_EPI = TCD_REGS_EPI, D_REGS_HELP = TCD_REGS_HELP;
static_assert(16 * D_REGS_ARITH + 4 * D_REGS_EPI + 4 * D_REGS_HELP <= 24 * D_REGS_LAUNCH && 24 * 32 * D_REGS_LAUNCH <= 65536, "register split");

template <int NL>
__global__ void __maxnreg__(D_REGS_LAUNCH) mvm_sym_tcd_kernel(const SymDArgs a) {
    if (!tcd_gate_open(a.gate, a.gate_max, a.gate_sum4_max)) return;   // coordinates too large for the cancellation in U: sym_tc5.cu's kernel takes the launch
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);      // (opaque: see sym_tc_dev.cuh opaque_tid)
    unsigned char* sm = static_cast<unsigned char*>(__cvta_shared_to_generic((size_t)base));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + d_bar(NL));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + d_bar(NL) + 256);

    const int tid = opaque_tid(), warp = tid >> 5, lane = tid & 31;
    const int bx = (int)opaque_u32(blockIdx.x), by = (int)opaque_u32(blockIdx.y), bz = (int)opaque_u32(blockIdx.z);
    const uint32_t bar0 = base + d_bar(NL);      // shared-memory address of the barrier block
#ifdef TCD_DEBUG_STAMPS
    uint32_t* dbg_sm = reinterpret_cast<uint32_t*>(sm + d_bar(NL) + 512);
    const bool stamping = a.dbg != nullptr && bx == 7 && by == 0 && bz == 0;
    if (stamping)
        for (int q = tid; q < DBG_NT * DBG_NC; q += D_THREADS) dbg_sm[q] = 0u;
#endif
    const int chunk = bz;
    Tile5Iter it;
    it.I = a.rb_begin + bx;
    it.B = a.nblocks;
    it.n = a.n;
    const int per = (a.half + a.nsplits - 1) / a.nsplits;
    it.k_begin = by * per;
    it.ntiles = 4 * (min(a.half, it.k_begin + per) - it.k_begin);
    if (it.ntiles < 0) it.ntiles = 0;
    int G = a.G, NB = a.NB;
    const int KS = a.KS;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < D_ZST; ++s) {
            mbar_init(&bars[BD_ZFULL + s], 1);
            mbar_init(&bars[BD_ZFREE + s], 1);
        }
#pragma unroll
        for (int s = 0; s < D_BST; ++s) mbar_init(&bars[BD_BFULL + s], 1);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bars[BD_SFULL + b], D_AW / 2);
            mbar_init(&bars[BD_TDONE + b], 1);
            mbar_init(&bars[BD_EREAD + b], 4);
            mbar_init(&bars[BD_D1EMPTY + b], D_AW);
        }
        mbar_init(&bars[BD_BCFULL], 1);
        mbar_init(&bars[BD_AFULL], 4);
#pragma unroll
        for (int s = 0; s < 2 * D_NBUF; ++s) {
            mbar_init(&bars[BD_D0FULL + s], 1);
            mbar_init(&bars[BD_D0FREE + s], D_AW / 2);
        }
        mbar_fence_init();
        // loop parameters, to be read back as register values (see below)
        uint32_t* pb = reinterpret_cast<uint32_t*>(sm + d_bar(NL) + 264);
        pb[0] = (uint32_t)G; pb[1] = (uint32_t)NB; pb[2] = (uint32_t)it.B; pb[3] = (uint32_t)it.I; pb[4] = (uint32_t)it.k_begin; pb[5] = (uint32_t)it.ntiles;
    }
    if (warp == D_AW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;
    // the parameters the loops test on every tile / batch come back from shared memory through volatile loads: ptxas reloads plain
    // kernel parameters from the constant bank at every use, and an LDC followed at once by the compare and branch that need it costs
    // its full latency -- the arithmetic warps' tile top had four of those in a row (profiles/tcd_timeline_r02.txt)
    {
        const uint32_t pa = bar0 + 264u;
        G = (int)lds5_volatile(pa); NB = (int)lds5_volatile(pa + 4u); it.B = (int)lds5_volatile(pa + 8u); it.I = (int)lds5_volatile(pa + 12u);
        it.k_begin = (int)lds5_volatile(pa + 16u); it.ntiles = (int)lds5_volatile(pa + 20u);
    }

    if (warp < D_AW) {
        // =========================================== arithmetic warps ===================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(D_REGS_ARITH));
        // two teams of eight warps take the tiles in turn (team = tile parity = S buffer), so that one team's exponentials fill the
        // XU pipe while the other waits for its exponents, loads them, or splits and stores its S tile.  Inside a team: TMEM lane
        // quadrant = warp & 3 (thread = row), half = which 16 of the tile's 32 columns.
        const int team = warp >> 3, half = (warp >> 2) & 1, part = 2 * team + half;
        const int rtid = tid & (T5_ROWS - 1);
        const long long row = (long long)it.I * T5_ROWS + rtid;
        const bool valid = row < a.n;
        const uint32_t lanes = (uint32_t)((warp & 3) * 32) << 16;  // this warp's TMEM lane quadrant
        f32x2 acc[D_FC / 2], comp[D_FC / 2];
#pragma unroll
        for (int q = 0; q < D_FC / 2; ++q) { acc[q] = 0ull; comp[q] = 0ull; }

        // the A operand of the distance MMAs goes to tensor memory once per CTA: warps 0..3 (thread = row = TMEM lane) read their
        // row's 128-byte lines of the pre-pass image (un-swizzling the 16-byte units) and tcgen05.st them, 32 columns at a time
        if (warp < 4 && it.next_live(0) < it.ntiles) {
            const unsigned char* arow = a.aimg + ((size_t)chunk * a.nblocks + it.I) * NL * 32768 + (size_t)rtid * 128;
#pragma unroll
            for (int lp = 0; lp < 2 * NL; ++lp) {      // lp = 2 * line + plane (tf32 part | remainder)
                uint32_t v[32];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint4 q = __ldg(reinterpret_cast<const uint4*>(arow + (size_t)lp * 16384 + (size_t)((u ^ (rtid & 7)) << 4)));
                    v[4 * u] = q.x; v[4 * u + 1] = q.y; v[4 * u + 2] = q.z; v[4 * u + 3] = q.w;
                }
                const uint32_t ta = tmem + D_TM_A + (uint32_t)((lp & 1) * NL * 32 + (lp >> 1) * 32) + lanes;
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
                    "%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(ta),
                    "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                    "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
                    "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                    : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar5_arrive_a(bar0 + 8u * (uint32_t)(BD_AFULL));
        }

        // every warp folds its rows' share (4 of the 16 right-hand sides) of a closed row-side epoch into the running total
        auto fold_epoch = [&](int e) {
            float d[D_FC], x[D_FC];
            const uint32_t ta = tmem + D_TM_D1 + 32u * (uint32_t)(e & 1) + (uint32_t)(part * D_FC) + lanes;
            tmem5_ld4(ta, d);            // Sh.Vh + Sl.Vh
            tmem5_ld4(ta + 16u, x);      // Sh.Vl
            tc5_fence_before();
            __syncwarp();
            if (elect_one()) mbar5_arrive_a(bar0 + 8u * (uint32_t)(BD_D1EMPTY + (e & 1)));
#pragma unroll
            for (int q = 0; q < D_FC / 2; ++q) {
                const f32x2 y = sub2(pack2(d[2 * q] + x[2 * q], d[2 * q + 1] + x[2 * q + 1]), comp[q]);
                const f32x2 tsum = add2(acc[q], y);
                comp[q] = sub2(sub2(tsum, acc[q]), y);
                acc[q] = tsum;
            }
        };

        int folded = 0;              // row-side epochs folded so far
        // what follows the exponentials of a tile: wait for the S buffer, fold closed row-side epochs, split and store S, hand it to
        // the S-side issuer.  (Running this under the first batch of the team's NEXT tile -- sums parked in 16 more registers -- was
        // measured 15-50 % slower: the two teams run in step, so every warp is in its tail at the same time either way, and the
        // extra registers spill.  profiles/tcd_variants_r02.txt)
        auto finish_tile = [&](int pj, const float* sp) {
            if (pj >= 2) {   // only now is the S buffer needed: tile pj-2 has left the tensor core (its exponentials ran under the S-side
                            // MMAs of the team's previous tile), and -- the S-side issuer commits in tile order -- every row-side epoch
                            // that ended at or before tile pj-2 is closed
                mbar5_wait_a(bar0 + 8u * (uint32_t)(BD_TDONE + team), (uint32_t)(((pj >> 1) - 1) & 1));
                tc5_fence_after();
                if (tid == 0 || tid == 256) TCD_STAMP(pj, 9);
                while ((folded + 1) * D_F <= pj - 1) fold_epoch(folded++);
            }
            if (tid == 0 || tid == 256) TCD_STAMP(pj, 10);
            // split, SWIZZLE_128B_BASE32B (32-byte chunks ^ row % 4), row-local stores (padding rows / columns hold exact zeros:
            // their exponent is >= D_PAD).  In row order a quarter-warp's eight 16-byte stores collide two by two (the swizzle only
            // spreads 4 rows); letting rows 4..7 of every eight store the halves of a 32-byte chunk in the other order removes the
            // conflicts at the price of 16 selects.  That pays when a tile is little more than its stores (one or two groups per
            // chunk: -9 %) and costs 1-2 % otherwise (profiles/tcd_timeline_r02.txt), hence the two forms.
            const uint32_t sc = base + D_S + (uint32_t)team * 32768u;
            if (G <= 2) {
                const bool sw = ((uint32_t)rtid >> 2) & 1u;
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        float x[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[e] = (sw != (i == 1)) ? sp[8 * pp + 4 + e] : sp[8 * pp + e];
                        float4 h, l;
                        h.x = tf32_hi5(x[0]); h.y = tf32_hi5(x[1]); h.z = tf32_hi5(x[2]); h.w = tf32_hi5(x[3]);
                        l.x = x[0] - h.x; l.y = x[1] - h.y; l.z = x[2] - h.z; l.w = x[3] - h.w;
                        const uint32_t q1 = (uint32_t)(half * 2 + pp), q0 = (uint32_t)i ^ (uint32_t)sw;      // piece q = 2 q1 + q0
                        const uint32_t off = (uint32_t)rtid * 128u + (((q1 ^ ((uint32_t)rtid & 3u)) << 5) | (q0 << 4));
                        sts5_v4(sc + off, h);
                        sts5_v4(sc + 16384u + off, l);
                    }
                }
            } else {
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const uint32_t q = (uint32_t)(half * 4 + qq);
                    float4 h, l;
                    h.x = tf32_hi5(sp[4 * qq]); h.y = tf32_hi5(sp[4 * qq + 1]); h.z = tf32_hi5(sp[4 * qq + 2]); h.w = tf32_hi5(sp[4 * qq + 3]);
                    l.x = sp[4 * qq] - h.x; l.y = sp[4 * qq + 1] - h.y; l.z = sp[4 * qq + 2] - h.z; l.w = sp[4 * qq + 3] - h.w;
                    const uint32_t off = (uint32_t)rtid * 128u + ((((q >> 1) ^ ((uint32_t)rtid & 3u)) << 5) | ((q & 1u) << 4));
                    sts5_v4(sc + off, h);
                    sts5_v4(sc + 16384u + off, l);
                }
            }
            if (tid == 0 || tid == 256) TCD_STAMP(pj, 11);
            if (!TCD_FENCE_ISSUER) fence5_async_smem();
            if (tid == 0 || tid == 256) TCD_STAMP(pj, 12);
            __syncwarp();
            if (elect_one()) mbar5_arrive_a(bar0 + 8u * (uint32_t)(BD_SFULL + team));
            if (tid == 0 || tid == 256) TCD_STAMP(pj, 13);
        };

        // D0 is handed over in batches of two groups: the team's item i = (tile, batch) lives in the team's D0 buffer i % 2 and is
        // released as soon as its exponents are in registers.  The hand-off sits on every warp's critical path (the teams run in
        // step, nothing hides it: 25 % of the arithmetic warps' time in profiles/ncu_r02b_tcd_cfg5b_summary.md), so its state is a
        // handful of precomputed addresses that toggle, and the diagonal block's exact-exponent fix lives in a second copy of the loop.
        static_assert(D_NBUF == 2, "the toggling below assumes two D0 buffers per team");
        const uint32_t tteam = tmem + D_TM_D0 + (uint32_t)(half * 16) + lanes + 64u * (uint32_t)(D_NBUF * team);   // + 64 ibuf: this thread's 16 columns
        const uint32_t full0 = bar0 + 8u * (uint32_t)(BD_D0FULL + D_NBUF * team), free0 = bar0 + 8u * (uint32_t)(BD_D0FREE + D_NBUF * team);
        uint32_t ibuf = 0, par = 0;      // buffer of the next batch, parity of its D0FULL phase
        float s[16];
        int dcol = -1;                   // diagonal block: the column of this thread's 16 that is the pair (row, row)
        int jcur = 0;
        // one batch = two groups (the last one of a chunk with an odd number of groups: one): both tcgen05.ld and their wait are ONE
        // asm statement with plain outputs, so that the exponents go from the load's destination registers straight into MUFU.EX2
        auto one_batch = [&](auto diag_tag, auto two_tag, int k) {
            constexpr bool DIAG = decltype(diag_tag)::value, TWO = decltype(two_tag)::value;
            mbar5_wait_a(full0 + 8u * ibuf, par);
            if (tid == 0 || tid == 256) TCD_STAMP(jcur, 1 + (k & 3));
            tc5_fence_after();
            const uint32_t ta = tteam + 64u * ibuf;
            uint32_t w[32];
            if constexpr (TWO) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%33];\n\t"
                    "tcgen05.wait::ld.sync.aligned;"
                    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
                      "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]),
                      "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]),
                      "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
                    : "r"(ta), "r"(ta + 32u)
                    : "memory");
            } else {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
                    "tcgen05.wait::ld.sync.aligned;"
                    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
                      "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                    : "r"(ta)
                    : "memory");
            }
            tc5_fence_before();
            __syncwarp();
            if (elect_one()) mbar5_arrive_a(free0 + 8u * ibuf);
            par ^= ibuf;            // the phase advances when the buffer index wraps (1 -> 0)
            ibuf ^= 1u;
            if constexpr (DIAG) {   // a pair with itself: the exact exponent (no cancellation error on the dominant entries of K)
#pragma unroll
                for (int gb = 0; gb < (TWO ? 2 : 1); ++gb) {
                    const int jg = chunk * G + 2 * k + gb;
                    const float nl = jg < a.J ? __ldg(a.nlc + jg) : D_PAD;
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (c == dcol) w[16 * gb + c] = __float_as_uint(nl);
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) s[c] += ex2_ftz(-__uint_as_float(w[c]));
            if constexpr (TWO) {
#pragma unroll
                for (int c = 0; c < 16; ++c) s[c] += ex2_ftz(-__uint_as_float(w[16 + c]));
            }
            if (tid == 0 || tid == 256) TCD_STAMP(jcur, 5 + (k & 3));
        };
        const int nfull = G >> 1;        // batches of two groups; an odd G leaves one batch of one group
        auto run_batches = [&](auto diag_tag) {
#pragma unroll 1
            for (int k = 0; k < nfull; ++k) one_batch(diag_tag, std::true_type{}, k);
            if (G & 1) one_batch(diag_tag, std::false_type{}, nfull);
        };

        int j = 0;                   // j counts the live tiles of BOTH teams
        LiveRun run = it.first_run();
        for (; run.t < it.ntiles; it.advance(run), ++j) {
            if ((j & 1) != team) continue;
            const int t = run.t;
            const bool diag = it.diag(t);
            jcur = j;
            if (tid == 0 || tid == 256) TCD_STAMP(j, 0);
#pragma unroll
            for (int c = 0; c < 16; ++c) s[c] = 0.f;
            if (diag) {
                dcol = (int)(row - (it.col0(t) + half * 16));
                run_batches(std::true_type{});
            } else {
                run_batches(std::false_type{});
            }
            finish_tile(j, s);
        }
        if (j > 0) {    // j = number of live tiles; the row issuer's last commit covers every earlier row-side MMA (the last two
                        // tiles in order: a parity wait must not fall more than one phase behind its barrier)
            if (j >= 2) mbar5_wait_a(bar0 + 8u * (uint32_t)(BD_TDONE + ((j - 2) & 1)), (uint32_t)(((j - 2) >> 1) & 1));
            mbar5_wait_a(bar0 + 8u * (uint32_t)(BD_TDONE + ((j - 1) & 1)), (uint32_t)(((j - 1) >> 1) & 1));
            tc5_fence_after();
            const int epochs = (j + D_F - 1) / D_F;
            while (folded < epochs) fold_epoch(folded++);
            if (valid) {
                double* dst = a.acc + row * T5_N + part * D_FC;
#pragma unroll
                for (int q = 0; q < D_FC / 2; ++q) {
                    float x, y;
                    unpack2(acc[q], x, y);
                    atomicAdd(dst + 2 * q, (double)x);
                    atomicAdd(dst + 2 * q + 1, (double)y);
                }
            }
        }
    } else if (warp < D_AW + 4) {
        // =========================================== epilogue warps (column side) ========================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(D_REGS_EPI));
        const int qd = warp - D_AW;                                // TMEM lane quadrant
        int j = 0, jc = 0;
        for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++j) {
            const int b = j & 1;
            const bool diag = it.diag(t);
            mbar_wait_ns(a.sleep_ns, &bars[BD_TDONE + b], (uint32_t)((j >> 1) & 1));
            tc5_fence_after();
            if (qd == 0 && lane == 0) TCD_STAMP(j, 21);
            // quadrants 0,1 hold the tf32-part rows of columns 0..15 / 16..31 (lanes 0..15), quadrants 2,3 the remainder rows
            float* P = reinterpret_cast<float*>(sm + D_EPI) + (jc & 1) * 1024;
            if (!diag) {
                const uint32_t ta = tmem + D_TM_D2 + 32u * (uint32_t)b + ((uint32_t)(qd * 32) << 16);
                float d[8], x[8], d2[8], x2[8];     // right-hand sides 0..7 | 8..15, [Vh part | Vl part] summed
                if (!(TCD_DIAG & 32)) tmem5_ld8x4(ta, ta + 16u, ta + 8u, ta + 24u, d, x, d2, x2);
                if (lane < 16) {
                    float4* dst = reinterpret_cast<float4*>(P + qd * 256 + lane * 16);
                    dst[0] = make_float4(d[0] + x[0], d[1] + x[1], d[2] + x[2], d[3] + x[3]);
                    dst[1] = make_float4(d[4] + x[4], d[5] + x[5], d[6] + x[6], d[7] + x[7]);
                    dst[2] = make_float4(d2[0] + x2[0], d2[1] + x2[1], d2[2] + x2[2], d2[3] + x2[3]);
                    dst[3] = make_float4(d2[4] + x2[4], d2[5] + x2[5], d2[6] + x2[6], d2[7] + x2[7]);
                }
            }
            tc5_fence_before();
            __syncwarp();
            if (lane == 0) mbar5_arrive(&bars[BD_EREAD + b]);
            if (!diag) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int et = tid - 32 * D_AW, c = et & 15;
                const long long c0 = it.col0(t);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int rr = (et >> 4) + 8 * k;      // tile column 0..31 = output row c0 + rr
                    const float v = P[(rr >> 4) * 256 + (rr & 15) * 16 + c] + P[(2 + (rr >> 4)) * 256 + (rr & 15) * 16 + c];
                    if (c0 + rr < a.n && !(TCD_DIAG & 1)) atomicAdd(a.acc + (c0 + rr) * T5_N + c, (double)v);
                }
                ++jc;
            }
            if (qd == 0 && lane == 0) TCD_STAMP(j, 22);
        }
    } else {
        // =========================================== helper warps =========================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(D_REGS_HELP));
        // role 0: S-side MMAs (row side, then column side of every tile); role 1: bulk copies; roles 2, 3: distance MMAs of team 0 / 1.
        // Every lane follows the barriers; tcgen05.mma / commit / cp.async.bulk are issued by one elected lane (elect_one).
        const int role = warp - D_AW - 4;
        const bool has_tiles = it.next_live(0) < it.ntiles;
        if (has_tiles && role == 0) {
            constexpr uint32_t IDESC_ROW_N32 = idesc5_tf32(128, 2 * T5_N, 0, 0), IDESC_ROW_N16 = idesc5_tf32(128, T5_N, 0, 0);
            constexpr uint32_t IDESC_COL = idesc5_tf32(64, 2 * T5_N, 1, 0);
            mbar_wait_ns(a.sleep_ns, &bars[BD_BCFULL], 0u);
            int j = 0;
            for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++j) {
                const int b = j & 1, e = j / D_F, bs = j % D_BST;
                mbar_wait_ns(a.sleep_ns, &bars[BD_BFULL + bs], (uint32_t)((j / D_BST) & 1));
                if (j % D_F == 0 && e >= 2) mbar_wait_ns(a.sleep_ns, &bars[BD_D1EMPTY + (e & 1)], (uint32_t)(((e >> 1) - 1) & 1));
                if (j >= 2) mbar_wait_ns(a.sleep_ns, &bars[BD_EREAD + b], (uint32_t)(((j >> 1) - 1) & 1));        // D2[b] has been read
                mbar_wait_ns(a.sleep_ns, &bars[BD_SFULL + b], (uint32_t)((j >> 1) & 1));
                if (TCD_FENCE_ISSUER) fence5_async_smem();     // the team's S stores (acquired through SFULL) -> async proxy (tensor core)
                tc5_fence_after();
                if (lane == 0) TCD_STAMP(j, 19);
                const bool col_work = !it.diag(t) && !(TCD_DIAG & 2);      // (nothing to do for the column side on the diagonal block)
                if (elect_one()) {
                    const uint32_t sbuf = base + D_S + (uint32_t)b * 32768u;
                    if (!(TCD_DIAG & 4)) {      // D1 += Sh.[Vh|Vl] + Sl.Vh, four k-steps of 8 tile columns
                        const uint32_t d1 = tmem + D_TM_D1 + 32u * (uint32_t)(e & 1);
                        const uint64_t dA_h = smem_desc5(sbuf, 512, 512, LAYOUT5_SW128_BASE32B);
                        const uint64_t dA_l = smem_desc5(sbuf + 16384u, 512, 512, LAYOUT5_SW128_BASE32B);
                        const uint64_t dB = smem_desc5(base + D_BT + (uint32_t)bs * 4096u, 16, 1024, LAYOUT5_SW128);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            umma5(d1, dA_h + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC_ROW_N32, (j % D_F != 0 || ks > 0) ? 1u : 0u);
                            umma5(d1, dA_l + (uint64_t)(ks * 2), dB + (uint64_t)(ks * 2), IDESC_ROW_N16, 1u);
                        }
                    }
                    if (col_work) {             // D2 = [Sh ; Sl]^T . [Vh|Vl], sixteen k-steps of 8 tile rows
                        const uint32_t d2 = tmem + D_TM_D2 + 32u * (uint32_t)b;
                        const uint64_t dA = smem_desc5(sbuf, 16384, 512, LAYOUT5_SW128_BASE32B);
                        const uint64_t dB = smem_desc5(base + D_BC, 16, 1024, LAYOUT5_SW128);
#pragma unroll
                        for (int g = 0; g < 16; ++g)
                            umma5(d2, dA + (uint64_t)((g * 1024) >> 4), dB + (uint64_t)(((g >> 2) * 4096 + (g & 3) * 32) >> 4), IDESC_COL, g > 0 ? 1u : 0u);
                    }
                    umma5_commit(&bars[BD_TDONE + b]);
                }
                __syncwarp();
                if (lane == 0) TCD_STAMP(j, 20);
            }
        } else if (has_tiles && role == 1) {
            // B images D_ZST deep (refilled at ZFREE of the tile that held the stage), B tiles of V D_BST deep (refilled at TDONE)
            const unsigned char* bsplit = reinterpret_cast<const unsigned char*>(a.bsplit);
            const unsigned char* bimg = a.bimg + (size_t)chunk * a.nblocks * 4 * NL * 8192;
            int tz = it.next_live(0), jz = 0, tb = tz, jb = 0;
            auto load_z = [&]() {
                const long long c0 = it.col0(tz);
                const int zs = jz % D_ZST;
                if (elect_one()) {
                    mbar_expect_tx(&bars[BD_ZFULL + zs], (uint32_t)NL * 8192u);
                    bulk_g2s(sm + d_bimg(NL) + (uint32_t)zs * NL * 8192u, bimg + (size_t)(c0 / T5_BN) * NL * 8192, (uint32_t)NL * 8192u, &bars[BD_ZFULL + zs]);
                }
                __syncwarp();
                tz = it.next_live(tz + 1);
                ++jz;
            };
            auto load_b = [&]() {
                const long long c0 = it.col0(tb);
                const int bs = jb % D_BST;
                if (elect_one()) {
                    mbar_expect_tx(&bars[BD_BFULL + bs], 4096u);
                    bulk_g2s(sm + D_BT + (uint32_t)bs * 4096u, bsplit + (c0 / T5_BN) * 4096, 4096u, &bars[BD_BFULL + bs]);
                }
                __syncwarp();
                tb = it.next_live(tb + 1);
                ++jb;
            };
            if (elect_one()) {
                mbar_expect_tx(&bars[BD_BCFULL], 16384u);
                bulk_g2s(sm + D_BC, bsplit + (long long)it.I * 16384, 16384u, &bars[BD_BCFULL]);
            }
            __syncwarp();
            for (int s = 0; s < D_ZST && tz < it.ntiles; ++s) load_z();
            for (int s = 0; s < D_BST && tb < it.ntiles; ++s) load_b();
            // the stage of tile j's B image is free once its distance MMAs have completed (ZFREE), the stage of its V tile once the
            // tile is through the S-side MMAs (TDONE); both barriers are followed in tile order
            int j = 0;
            for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++j) {
                mbar_wait_ns(a.sleep_ns, &bars[BD_ZFREE + j % D_ZST], (uint32_t)((j / D_ZST) & 1));
                if (tz < it.ntiles) load_z();
                mbar_wait_ns(a.sleep_ns, &bars[BD_TDONE + (j & 1)], (uint32_t)((j >> 1) & 1));
                if (tb < it.ntiles) load_b();
            }
        } else if (has_tiles && role - 2 < D_NDI) {
            // issuer w owns team w's tiles (tile parity): U = Ah.Bh + Al.Bh + Ah.Bl for both groups of every batch into the batch's D0
            // buffer, KS k-steps each.  It follows the B-image barriers of ALL tiles (a parity wait must not skip a phase).
            const int w = role - 2;
            constexpr uint32_t IDESC_D0 = idesc5_tf32(128, T5_BN, 0, 0);
            mbar_wait_ns(a.sleep_ns, &bars[BD_AFULL], 0u);
            tc5_fence_after();
            int jd = 0;
            uint32_t ibuf = 0, iuse = 0;       // this team's (tile, batch) counter modulo / divided by D_NBUF
            for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), ++jd) {
                const int zs = jd % D_ZST;
                if (D_ZST % 2 == 0 && (jd & 1) != w) continue;      // (even stage count: the other team's tiles never use this team's stages)
                mbar_wait_ns(a.sleep_ns, &bars[BD_ZFULL + zs], (uint32_t)((jd / D_ZST) & 1));
                if ((jd & 1) != w) continue;
                tc5_fence_after();
                if (lane == 0) TCD_STAMP(jd, 18);
                const uint32_t bst = base + d_bimg(NL) + (uint32_t)zs * NL * 8192u;
                for (int k = 0; k < NB; ++k) {
                    const uint32_t buf = (uint32_t)(D_NBUF * w) + ibuf;
                    if (iuse >= 1) {
                        mbar_wait_ns(a.sleep_ns, &bars[BD_D0FREE + buf], (iuse - 1) & 1u);
                        tc5_fence_after();
                    }
                    if (elect_one()) {
#pragma unroll
                        for (int gb = 0; gb < 2; ++gb) {
                            const int g = 2 * k + gb;
                            if (g < G) {
                                const uint32_t d0 = tmem + D_TM_D0 + 64u * buf + 32u * (uint32_t)gb;
                                for (int ks = 0; ks < ((TCD_DIAG & 8) ? 0 : KS); ++ks) {
                                    const int ksi = g * KS + ks;
                                    const uint32_t l = (uint32_t)(ksi >> 2);
                                    const uint64_t o = (uint64_t)((ksi & 3) * 2);
                                    const uint32_t aH = tmem + D_TM_A + 8u * (uint32_t)ksi, aL = aH + (uint32_t)(NL * 32);      // A from tensor memory
                                    const uint64_t dBh = smem_desc5(bst + l * 8192u, 16, 1024, LAYOUT5_SW128) + o;
                                    const uint64_t dBl = smem_desc5(bst + l * 8192u + 4096u, 16, 1024, LAYOUT5_SW128) + o;
                                    umma5_ts(d0, aH, dBh, IDESC_D0, ks > 0 ? 1u : 0u);
                                    umma5_ts(d0, aL, dBh, IDESC_D0, 1u);
                                    umma5_ts(d0, aH, dBl, IDESC_D0, 1u);
                                }
                            }
                        }
                        umma5_commit(&bars[BD_D0FULL + buf]);
                        if (k == NB - 1) umma5_commit(&bars[BD_ZFREE + zs]);     // the tile's B image has been consumed
                    }
                    __syncwarp();
                    if (lane == 0) TCD_STAMP(jd, 14 + (k & 3));
                    if (++ibuf == D_NBUF) { ibuf = 0; ++iuse; }
                }
            }
        }
        __syncwarp();
    }
    tc5_fence_before();
    __syncthreads();
    if (warp == D_AW) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
#ifdef TCD_DEBUG_STAMPS
    if (stamping)
        for (int q = tid; q < DBG_NT * DBG_NC; q += D_THREADS) reinterpret_cast<uint32_t*>(a.dbg)[q] = dbg_sm[q];
#endif
}

// ---- operand images ---------------------------------------------------------------------------------------------------------
// column sums of the packed coordinates (for centring): sums[chunk*CP + q] += sum over a slab of rows
__global__ void tcd_colsum_kernel(const float* __restrict__ zp, long long n, int CP, int nchunks, double* __restrict__ sums) {
    __shared__ double red[8][33];
    const int q = threadIdx.x, chunk = blockIdx.y;
    double acc = 0.0;
    if (q < CP)
        for (long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y; r < n; r += (long long)gridDim.x * blockDim.y)
            acc += (double)__ldg(zp + ((long long)chunk * n + r) * CP + q);
    red[threadIdx.y][q] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && q < CP) {
        for (int y = 1; y < 8; ++y) acc += red[y][q];
        atomicAdd(sums + chunk * CP + q, acc);
    }
}

// one thread per (row, chunk, k-step of a 128-byte line): 8 floats of the A line and 8 of the B line, split and swizzled
__global__ void tcd_build_images_kernel(const float* __restrict__ zp, long long n, long long n_pad, Layout lay, const float* __restrict__ nlc,
                                        const double* __restrict__ colsum, int GT, int KS, int NL, int nch, unsigned char* __restrict__ aimg,
                                        unsigned char* __restrict__ bimg, unsigned* __restrict__ gate) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nks = 4 * NL;
    const long long per_chunk = n_pad * nks;
    float maxsq = 0.f;
    if (idx < per_chunk * nch) {
        const int chunk = (int)(idx / per_chunk);
        const long long rem = idx - (long long)chunk * per_chunk;
        const int ksi = (int)(rem / n_pad);             // consecutive threads = consecutive rows
        const long long row = rem - (long long)ksi * n_pad;
        const int g = ksi / KS, ks = ksi - g * KS, jg = chunk * GT + g, K = lay.K;
        float A[8], B[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { A[e] = 0.f; B[e] = 0.f; }
        if (g < GT) {
            const bool live = row < n && jg < lay.J;
            if (live) {
                const int lc = jg / lay.G, off = (jg % lay.G) * lay.KP;      // the group's chunk and offset in the packed layout
                const float* src = zp + ((long long)lc * n + row) * lay.CP + off;
                const double* mu = colsum + lc * lay.CP + off;
                const double inv_n = 1.0 / (double)n;
                double sq = 0.0;
#pragma unroll
                for (int e = 0; e < 6; ++e) {
                    const int m = 6 * ks + e;
                    if (m < K) {
                        const float z = (float)((double)__ldg(src + m) - mu[m] * inv_n);
                        A[e] = -2.f * z;
                        B[e] = z;
                        sq += (double)z * (double)z;
                    }
                }
                const float sqf = (float)sq;
                A[6] = ks == 0 ? fminf(sqf + __ldg(nlc + jg), D_PAD) : sqf;
                B[6] = 1.f;
                A[7] = 1.f;
                B[7] = sqf;
                if (ks == 0) {      // the whole group's squared norm decides whether the tensor-core distances are accurate enough
                    double tot = 0.0;
                    for (int m = 0; m < K; ++m) { const double z = (double)__ldg(src + m) - mu[m] * inv_n; tot += z * z; }
                    maxsq = (float)tot;
                }
            } else if (ks == 0) {   // padding rows / columns / groups: exponent >= D_PAD, i.e. an exact zero
                A[6] = D_PAD; B[6] = 1.f; A[7] = 1.f; B[7] = D_PAD;
            }
        }
        float4 Ah[2], Al[2], Bh[2], Bl[2];
        float* pAh = reinterpret_cast<float*>(Ah); float* pAl = reinterpret_cast<float*>(Al);
        float* pBh = reinterpret_cast<float*>(Bh); float* pBl = reinterpret_cast<float*>(Bl);
#pragma unroll
        for (int e = 0; e < 8; ++e) {   // round-to-nearest split (a truncating split biases U by ~3e-7 |z|^2: tools/tcd_check.py adv)
            pAh[e] = tf32_rn(A[e]); pAl[e] = tf32_rn(A[e] - pAh[e]);
            pBh[e] = tf32_rn(B[e]); pBl[e] = tf32_rn(B[e] - pBh[e]);
        }
        const int l = ksi >> 2;
        const uint32_t c16 = (uint32_t)(ksi & 3) * 2u;
        const long long nblocks = n_pad / T5_ROWS;
        {
            const uint32_t r = (uint32_t)(row & 127);
            unsigned char* dst = aimg + (((size_t)chunk * nblocks + (row >> 7)) * NL + l) * 32768 + (size_t)r * 128;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint32_t off = ((c16 + hh) ^ (r & 7u)) << 4;
                *reinterpret_cast<float4*>(dst + off) = Ah[hh];
                *reinterpret_cast<float4*>(dst + 16384 + off) = Al[hh];
            }
        }
        {
            const uint32_t r = (uint32_t)(row & 31);
            unsigned char* dst = bimg + (((size_t)chunk * nblocks * 4 + (row >> 5)) * NL + l) * 8192 + (size_t)r * 128;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint32_t off = ((c16 + hh) ^ (r & 7u)) << 4;
                *reinterpret_cast<float4*>(dst + off) = Bh[hh];
                *reinterpret_cast<float4*>(dst + 4096 + off) = Bl[hh];
            }
        }
    }
    double sum4 = (double)maxsq * (double)maxsq;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        maxsq = fmaxf(maxsq, __shfl_xor_sync(0xffffffffu, maxsq, off));
        sum4 += __shfl_xor_sync(0xffffffffu, sum4, off);
    }
    if ((threadIdx.x & 31) == 0 && maxsq > 0.f) {
        atomicMax(gate, __float_as_uint(maxsq));
        atomicAdd(reinterpret_cast<double*>(gate + 2), sum4);
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
TcdPlan plan_tcd(const Layout& lay) {
    TcdPlan p;
    std::memset(&p, 0, sizeof(p));
    static const int enabled = [] { const char* e = getenv("RPGP_SYM_TCD"); return e ? atoi(e) : 1; }();
    static const int kmin = [] { const char* e = getenv("RPGP_SYM_TCD_KMIN"); return e ? atoi(e) : 4; }();
    if (!enabled || lay.base != 0 || lay.K < kmin || lay.K > 24) return p;      // (RBF only: the exponent IS the distance)
    p.KS = (lay.K + 5) / 6;       // six coordinates + their two partial norms per k-step of eight
    const int gmax = 8 / p.KS;
    p.nchunks = (lay.J + gmax - 1) / gmax;
    p.GT = (lay.J + p.nchunks - 1) / p.nchunks;
    p.NL = (p.GT * p.KS + 3) / 4;
    p.NB = (p.GT + 1) / 2;
    p.GB = 2;
    p.supported = 1;
    return p;
}

size_t tcd_workspace_bytes(long long n, const Layout& lay) {
    const TcdPlan p = plan_tcd(lay);
    if (!p.supported) return 0;
    const size_t nblocks = (size_t)((n + T5_ROWS - 1) / T5_ROWS);
    return D_WS_HEADER + 2 * (size_t)p.nchunks * nblocks * p.NL * 32768;
}

TcdGate tcd_gate(long long n, const Layout& lay) {
    // norm-wise error model (profiles/tcd_accuracy_r01.txt): a row whose centred, scaled squared group norm is r2 carries a relative
    // error ~ 2..3.5e-8 * r2 on its near pairs, so ||error|| / ||K.V|| <~ 3.5e-8 * sqrt(mean r2^2): bound the root mean square of the
    // squared norms, and cap single rows at 10x that
    TcdGate g;
    const float b = tcd_gate_bound(), cap = 10.f * b;
    std::memcpy(&g.max_bits, &cap, sizeof(float));
    g.sum4_max = (double)b * (double)b * (double)n * (double)lay.J;
    return g;
}

float tcd_gate_bound() {
    // largest scaled, centred squared group norm for which U from the tensor core keeps K.V inside 1e-5 even when every pair that
    // matters sits at that radius (measured: tools/tcd_check.py adv, profiles/tcd_accuracy_r01.txt)
    static const float bound = [] { const char* e = getenv("RPGP_SYM_TCD_AMAX"); return e ? (float)atof(e) : 200.f; }();
    return bound;
}

int launch_sym_tcd(const float* zp, long long n, const Layout& lay, const float* nlc, const float* bsplit, double* acc, int nblocks,
                   int rb_begin, int nrb, void* ws, size_t ws_bytes, const unsigned** gate_out, cudaStream_t st) {
    const TcdPlan p = plan_tcd(lay);
    if (!p.supported) return ERR_UNSUPPORTED;
    const size_t img_bytes = (size_t)p.nchunks * nblocks * p.NL * 32768;
    if (ws == nullptr || ws_bytes < D_WS_HEADER + 2 * img_bytes) {
        set_error("mvm_sym: distance-image workspace %zu bytes < required %zu", ws_bytes, D_WS_HEADER + 2 * img_bytes);
        return ERR_WORKSPACE;
    }
    unsigned* gate = (unsigned*)ws;
    double* colsum = (double*)((unsigned char*)ws + 1024);
    unsigned char* aimg = (unsigned char*)ws + D_WS_HEADER;
    unsigned char* bimg = aimg + img_bytes;
    if ((size_t)lay.nchunks * lay.CP * sizeof(double) > D_WS_HEADER - 1024) return ERR_UNSUPPORTED;
    RPGP_CUDA_OK(cudaMemsetAsync(gate, 0, D_WS_HEADER, st));
    tcd_colsum_kernel<<<dim3((unsigned)std::min<long long>((n + 7) / 8, 1184), (unsigned)lay.nchunks), dim3(32, 8), 0, st>>>(zp, n, lay.CP, lay.nchunks, colsum);
    note_launch();
    RPGP_CUDA_OK(cudaGetLastError());
    const long long n_pad = (long long)nblocks * T5_ROWS;
    const long long total = n_pad * 4 * p.NL * p.nchunks;
    tcd_build_images_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(zp, n, n_pad, lay, nlc, colsum, p.GT, p.KS, p.NL, p.nchunks, aimg, bimg,
                                                                             gate);
    note_launch();
    RPGP_CUDA_OK(cudaGetLastError());

    SymDArgs a;
    a.aimg = aimg; a.bimg = bimg; a.bsplit = bsplit; a.nlc = nlc; a.acc = acc; a.gate = gate;
    const TcdGate gb = tcd_gate(n, lay);
    a.gate_max = gb.max_bits;
    a.gate_sum4_max = gb.sum4_max;
    a.n = n; a.nblocks = nblocks; a.half = nblocks / 2 + 1; a.rb_begin = rb_begin;
    a.G = p.GT; a.KS = p.KS; a.NB = p.NB; a.J = lay.J;
    static const int sleep_env = [] { const char* e = getenv("RPGP_TCD_SLEEP"); return e ? atoi(e) : D_SLEEP; }();
    a.sleep_ns = sleep_env;
    a.dbg = nullptr;
#ifdef TCD_DEBUG_STAMPS
    static const int dbg_env = [] { const char* e = getenv("RPGP_TCD_DBG"); return e ? atoi(e) : 0; }();
    if (dbg_env) {
        RPGP_CUDA_OK(cudaMalloc(&a.dbg, DBG_NT * DBG_NC * sizeof(unsigned)));
        RPGP_CUDA_OK(cudaMemset(a.dbg, 0, DBG_NT * DBG_NC * sizeof(unsigned)));
    }
    const size_t dbg_smem = 4096;
#else
    const size_t dbg_smem = 0;
#endif
    static const int splits_env = [] { const char* e = getenv("RPGP_SYM_SPLITS"); return e ? atoi(e) : 0; }();
    long long want = pick_splits((long long)nrb * p.nchunks, a.half, 148);      // one CTA per SM
    if (splits_env > 0) want = splits_env;
    want = std::max<long long>(1, std::min<long long>(want, a.half));
    a.nsplits = (int)want;
    dim3 grid((unsigned)nrb, (unsigned)a.nsplits, (unsigned)p.nchunks);
    cudaError_t e;
    if (p.NL == 1) {
        e = cudaFuncSetAttribute(mvm_sym_tcd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(d_smem_bytes(1) + dbg_smem));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mvm_sym_tcd_kernel<1>)");
        mvm_sym_tcd_kernel<1><<<grid, D_THREADS, d_smem_bytes(1) + dbg_smem, st>>>(a);
    } else {
        e = cudaFuncSetAttribute(mvm_sym_tcd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(d_smem_bytes(2) + dbg_smem));
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(mvm_sym_tcd_kernel<2>)");
        mvm_sym_tcd_kernel<2><<<grid, D_THREADS, d_smem_bytes(2) + dbg_smem, st>>>(a);
    }
    note_launch();
    *gate_out = gate;
#ifdef TCD_DEBUG_STAMPS
    if (a.dbg) {   // debugging aid only: synchronises and prints the stamps relative to the first stamped tile's top
        static unsigned h[DBG_NT * DBG_NC];
        RPGP_CUDA_OK(cudaStreamSynchronize(st));
        RPGP_CUDA_OK(cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(a.dbg);
        fprintf(stderr, "clk relative to the first stamped tile's top; team = tile parity\n"
                        "tile |   top | D0FULL passed b0..b3      | exps issued b0..b3        |  TDONE   fold stores  fence arrive |"
                        " dist: ZFULL, issued b0..b3        | S: ready commit | epi: TDONE  done\n");
        const unsigned o = h[0];
        for (int r = 0; r < DBG_NT; ++r) {
            fprintf(stderr, "%3d  |", DBG_T0 + r);
            const int order[23] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 18, 14, 15, 16, 17, 19, 20, 21, 22};
            for (int q = 0; q < 23; ++q) {
                const unsigned v = h[r * DBG_NC + order[q]];
                if (v) fprintf(stderr, " %6d", (int)(v - o)); else fprintf(stderr, "     -1");
                if (q == 0 || q == 4 || q == 8 || q == 13 || q == 18 || q == 20) fprintf(stderr, " |");
            }
            fprintf(stderr, "\n");
        }
    }
#endif
    return cuda_fail(cudaGetLastError(), "mvm_sym_tcd_kernel launch");
}

}  // namespace rpgp

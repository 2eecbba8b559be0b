// sym_tc_dev.cuh -- device helpers shared by the symmetric tensor-core kernels (sym_tc5.cu: direct differences,
// sym_tcd.cu: distances on tcgen05): UMMA descriptors, tcgen05 issue / commit / fences, TMEM loads, the tf32 split,
// the SWIZZLE_128B address function and the cyclic enumeration of the unique 128-row block pairs.
#pragma once
#include "rpgp_common.cuh"

namespace rpgp {
namespace tcdev {

constexpr int T5_ROWS = 128;     // rows per CTA
constexpr int T5_BN = 32;        // columns per tile
constexpr int T5_N = 16;         // padded right-hand sides

constexpr uint32_t LAYOUT5_SW128 = 2, LAYOUT5_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t smem_desc5(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc5_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma5(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand in TENSOR MEMORY (lanes = rows of M = 128, one 32-bit column per k; k-step s of 8 at column 8 s): no shared-memory read
// for A, ~35 clk per M128 N32 K8 MMA instead of ~63 (tools/umma_probe4.cu, profiles/umma_rate_r02.txt)
__device__ __forceinline__ void umma5_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma5_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar5_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// helper warps poll with a back-off so that their spinning does not take issue slots from the arithmetic warps
template <int NS_SLEEP = 64>
__device__ __forceinline__ void mbar5_wait_sleep(uint64_t* bar, uint32_t parity) {
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
        __nanosleep(NS_SLEEP);
    }
}
// One lane of a converged warp (elect.sync).  Issuing tcgen05.mma / tcgen05.commit / cp.async.bulk under this predicate -- instead of
// under `lane == 0` -- lets ptxas keep the descriptors in uniform registers and emit the UTCHMMAs back to back: in code it must assume
// divergent it wraps EVERY uniform-datapath instruction in an ELECT / BRA.U.ANY waterfall loop whose branch waits for the
// instruction's scoreboard (the "~150 clk per MMA, per issuing warp" of profiles/umma_microbench_r01.txt; SASS in
// profiles/umma_issue_r02.txt).  Every lane of the warp must reach the call.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, px;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence5_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem5_ld16(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void tmem5_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(r[q]);
}
__device__ __forceinline__ float tf32_hi5(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__host__ __device__ __forceinline__ uint32_t sw128_5(uint32_t row, uint32_t kk) {
    return row * 128u + ((((kk >> 2) ^ (row & 7u)) << 4) | ((kk & 3u) << 2));
}

// threadIdx.x as a value ptxas cannot re-derive from the special register: under register pressure it rematerialises threadIdx-based
// values with S2R SR_TID inside the hot loops, and S2R queues behind the MUFU.EX2 stream of the sub-partition -- hundreds of clk
// each in these XU-bound kernels (profiles/tcd_timeline_r02.txt).  A shuffle from the own lane is the identity ptxas does not see
// through.  Call once, with the whole warp converged.
__device__ __forceinline__ int opaque_tid() {
    const int t = (int)threadIdx.x;
    return __shfl_sync(0xffffffffu, t, t & 31);
}
// the same for warp-uniform values (blockIdx, the shared-memory window: S2R SR_CTAID / SR_CgaCtaId; kernel parameters: LDC).  The
// shuffle is written in PTX: the compiler folds __shfl_sync of a value it knows to be uniform.
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("shfl.sync.idx.b32 %0, %1, 0, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
// barrier operations on 32-bit shared-memory addresses (kept in a register from an opaque base instead of re-derived from a pointer)
__device__ __forceinline__ void mbar5_wait_a(uint32_t addr, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_A:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE_A;\n"
        "bra WAIT_LOOP_A;\n"
        "WAIT_DONE_A:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}
// a value read back from shared memory with a volatile load: the only form of a kernel parameter that ptxas will not reload from
// the constant bank (it folds even a shuffle of a uniform value)
__device__ __forceinline__ uint32_t lds5_volatile(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void mbar5_arrive_a(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void sts5_v4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// tile enumeration shared by all roles: offsets k in [k_begin, k_end), four 32-column tiles per 128-column block
struct LiveRun { int t, lim; };      // [t, lim): consecutive live tiles inside one 128-column block (t == ntiles: no tile left)
struct Tile5Iter {
    int I, B, k_begin, ntiles;
    long long n;
    __host__ __device__ __forceinline__ int block_of(int k) const { int Ip = I + k; return Ip >= B ? Ip - B : Ip; }
    __host__ __device__ __forceinline__ bool offset_active(int k) const { return !((B % 2 == 0) && (k == B / 2) && (I >= B / 2)); }
    __host__ __device__ __forceinline__ long long col0(int t) const { return (long long)block_of(k_begin + (t >> 2)) * T5_ROWS + (t & 3) * T5_BN; }
    __host__ __device__ __forceinline__ bool live(int t) const { return offset_active(k_begin + (t >> 2)) && col0(t) < n; }
    // (offsets k stay below B: half = B / 2 + 1 <= B, so block_of(k) == I exactly when k == 0)
    __host__ __device__ __forceinline__ bool diag(int t) const { return k_begin + (t >> 2) == 0; }
    __host__ __device__ __forceinline__ int next_live(int t) const { while (t < ntiles && !live(t)) ++t; return t; }
    // the same enumeration with the liveness test paid once per block instead of once per tile (a tile's neighbours in its block
    // are live up to the first column >= n): advance() is an increment and a compare for three tiles out of four
    __host__ __device__ __forceinline__ LiveRun run_from(int t) const {
        LiveRun r;
        r.t = next_live(t);
        r.lim = r.t;
        if (r.t < ntiles) {
            const int blk_end = min(ntiles, (r.t | 3) + 1);
            r.lim = r.t + 1;
            while (r.lim < blk_end && live(r.lim)) ++r.lim;
        }
        return r;
    }
    __host__ __device__ __forceinline__ LiveRun first_run() const { return run_from(0); }
    __host__ __device__ __forceinline__ void advance(LiveRun& r) const {
        if (++r.t >= r.lim) r = run_from(r.t);
    }
};

// gate of the distance-on-tensor-core path: words written by its pre-pass = {bits of max r2, -, double sum of r2^2} where r2 is
// the centred, scaled squared norm of a (row, group).  Both symmetric kernels evaluate it: exactly one of them runs.
__device__ __forceinline__ bool tcd_gate_open(const unsigned* gate, unsigned max_bits, double sum4_max) {
    return gate[0] <= max_bits && *reinterpret_cast<const double*>(gate + 2) <= sum4_max;
}

}  // namespace tcdev
}  // namespace rpgp

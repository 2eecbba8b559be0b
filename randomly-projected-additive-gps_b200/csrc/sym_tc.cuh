// sym_tc.cuh -- launchers of the symmetric tensor-core K(Z,Z).V kernels (sym_tc5.cu: direct differences, every layout;
// sym_tcd.cu: K > 1 groups with the squared distances on tcgen05; earlier variants live in experiments/)
#pragma once
#include "rpgp_common.cuh"

namespace rpgp {
struct Layout;

// chunking of the distance-on-tensor-core path: GT groups of KS k-steps (8 tf32 each) per chunk, NL 128-byte lines per row
struct TcdPlan {
    int supported, GT, KS, NL, NB, GB, nchunks;
};
TcdPlan plan_tcd(const Layout& lay);
size_t tcd_workspace_bytes(long long n, const Layout& lay);   // gate word + A / B operand images (0 when unsupported)
float tcd_gate_bound();
struct TcdGate {
    unsigned max_bits;   // bits of the largest admissible squared group norm of a single row
    double sum4_max;     // largest admissible sum over (row, group) of the squared squared norms
};
TcdGate tcd_gate(long long n, const Layout& lay);
// builds the operand images and launches the kernel on the unique block pairs of row blocks [rb_begin, rb_begin + nrb);
// *gate_out points at the device word the direct-difference kernel must consult (it runs only when the word exceeds the bound)
int launch_sym_tcd(const float* zp, long long n, const Layout& lay, const float* nlc, const float* bsplit, double* acc, int nblocks,
                   int rb_begin, int nrb, void* ws, size_t ws_bytes, const unsigned** gate_out, cudaStream_t st);

// Column splits of a launch of `items` = row blocks x coordinate chunks CTAs-per-split on `slots` resident CTA slots (148 SMs x CTAs
// per SM): every CTA of a split does the same work, so the launch runs in waves and the last one should be full.  The smallest split
// count whose wave efficiency waves / ceil(waves) reaches 98.5 % (at least eight waves), else the best one; at most `max_splits`
// (= the number of block offsets).  Round 1 took ceil(16 slots / items): 26.4 waves at n = 1M on one GPU (2.3 % tail), 16.5 on eight.
inline int pick_splits(long long items, long long max_splits, int slots) {
    int best = 1;
    double best_eff = 0.0;
    const long long cap = max_splits < 64 ? max_splits : 64;
    for (long long s = 1; s <= cap; ++s) {
        const double waves = (double)(items * s) / slots;
        const double full = (double)(long long)(waves + 0.999999);
        // the last split of a row block may be shorter: per = ceil(max_splits / s) offsets, the others carry the imbalance
        const long long per = (max_splits + s - 1) / s;
        const double balance = (double)max_splits / (double)(per * s);
        const double eff = waves / full * balance;
        if (waves >= 8.0 && eff >= 0.985) return (int)s;
        if (eff > best_eff + 1e-9) { best_eff = eff; best = (int)s; }
    }
    return best;
}

// FP64 accumulators [n][16] (rounded up to 1 KB) + the pre-split right-hand sides (16 KB per 128-row block) + slack
inline size_t sym_base_workspace_bytes(long long n) {
    return (((size_t)n * 16 * sizeof(double) + 1023) & ~(size_t)1023) + (size_t)((n + 127) / 128) * 16384 + 1024;
}
inline size_t sym_workspace_bytes(long long n, const Layout& lay) { return sym_base_workspace_bytes(n) + tcd_workspace_bytes(n, lay); }
// both uses of every kernel value on the tensor cores: the default for square products
int launch_sym_tc5(const float* zp, long long n, const Layout& lay, const float* nlc, const float* V16, int t, float* out, int ldo,
                   int rb_begin, int rb_end, void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace rpgp

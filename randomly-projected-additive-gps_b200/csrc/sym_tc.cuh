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

// FP64 accumulators [n][16] (rounded up to 1 KB) + the pre-split right-hand sides (16 KB per 128-row block) + slack
inline size_t sym_base_workspace_bytes(long long n) {
    return (((size_t)n * 16 * sizeof(double) + 1023) & ~(size_t)1023) + (size_t)((n + 127) / 128) * 16384 + 1024;
}
inline size_t sym_workspace_bytes(long long n, const Layout& lay) { return sym_base_workspace_bytes(n) + tcd_workspace_bytes(n, lay); }
// both uses of every kernel value on the tensor cores: the default for square products
int launch_sym_tc5(const float* zp, long long n, const Layout& lay, const float* nlc, const float* V16, int t, float* out, int ldo,
                   int rb_begin, int rb_end, void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace rpgp

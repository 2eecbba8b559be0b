// sym_tc.cuh -- launcher of the symmetric tensor-core K(Z,Z).V kernel (sym_tc5.cu; earlier variants live in experiments/)
#pragma once
#include "rpgp_common.cuh"

namespace rpgp {
// FP64 accumulators [n][16] (rounded up to 1 KB) + the pre-split right-hand sides (16 KB per 128-row block) + slack
inline size_t sym_workspace_bytes(long long n) {
    return (((size_t)n * 16 * sizeof(double) + 1023) & ~(size_t)1023) + (size_t)((n + 127) / 128) * 16384 + 512;
}
// both uses of every kernel value on the tensor cores (sym_tc5.cu): the default
struct Layout;
int launch_sym_tc5(const float* zp, long long n, const Layout& lay, const float* nlc, const float* V16, int t, float* out, int ldo,
                   int rb_begin, int rb_end, void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace rpgp

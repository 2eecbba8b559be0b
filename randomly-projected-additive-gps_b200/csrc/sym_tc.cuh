// sym_tc.cuh -- launcher of the symmetric tensor-core K(Z,Z).V kernel (sym_tc3.cu; earlier variants live in experiments/)
#pragma once
#include "rpgp_common.cuh"

namespace rpgp {
// FP64 accumulators [n][16] (+ slack) for the symmetric kernel
inline size_t sym_workspace_bytes(long long n) { return (size_t)n * 16 * sizeof(double) + 512; }
// warp-specialised variant (sym_tc3.cu): row side in registers, column side on the tensor cores issued by dedicated warps
int launch_sym_tc3(const float* zp, long long n, int CP, const float* nlc, const float* V16, int t, float* out, int ldo,
                   int rb_begin, int rb_end, void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace rpgp

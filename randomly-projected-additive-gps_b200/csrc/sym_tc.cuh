// sym_tc.cuh -- launcher of the symmetric tensor-core K(Z,Z).V kernel (sym_tc.cu)
#pragma once
#include "rpgp_common.cuh"

namespace rpgp {
size_t sym_workspace_bytes(long long n);
// out (n x ldo) = contributions of the unique block pairs {I, I'} owned by row blocks [rb_begin, rb_end) to K(Z,Z).V
// (all of K.V when the range is [0, ceil(n/128))).  V must be padded to [n][16].  finalize=0 leaves the FP64
// accumulators in the workspace.
int launch_sym_tc(const float* zp, long long n, int CP, const float* nlc, const float* V, int ldv, int t, float* out, int ldo,
                  int rb_begin, int rb_end, int finalize, void* workspace, size_t workspace_bytes, cudaStream_t st);
// warp-specialised variant (sym_tc3.cu): row side in registers, column side on the tensor cores issued by dedicated warps
int launch_sym_tc3(const float* zp, long long n, int CP, const float* nlc, const float* V16, int t, float* out, int ldo,
                   int rb_begin, int rb_end, void* workspace, size_t workspace_bytes, cudaStream_t st);
}  // namespace rpgp

// aux_kernels.cuh -- launchers of the auxiliary kernels (aux_kernels.cu) and the layout descriptor shared with api.cu.
#pragma once
#include "rpgp_common.cuh"

namespace rpgp {

struct Layout {  // binary-compatible with rpgp_layout in include/rpgp.h
    int J, K, CP, nchunks, KP, G;
    int base;   // base kernel of every group: 0 RBF, 1 Matern-1.5, 2 inverse multiquadric (kv_kernels.cuh)
};

int launch_pack_coords(const float* Z, long long n, long long ld, const Layout& lay, float scale, float* Zp, cudaStream_t st);
int launch_pack_log2c(const float* c, const Layout& lay, float* nlc, cudaStream_t st);
// Zp: packed planes (or NULL), Zn: natural n x J*K rows, row stride ldz, NOT multiplied by `scale` (or NULL)
int launch_project(const float* X, long long n, int d, long long ldx, const float* W, const float* pre_inv,
                   const float* post_inv, const Layout& lay, float scale, float* Zp, float* Zn, long long ldz, cudaStream_t st);
// project_tc.cu: the same on tcgen05 (3xTF32), and the vector-Jacobian product dW[q][k] = sum_i dZ[i][q] X[i][k]
bool project_tc_supported(int d, const Layout& lay);
int launch_project_tc(const float* X, long long n, int d, long long ldx, const float* W, const float* pre_inv, const float* post_inv,
                      const Layout& lay, float scale, float* Zp, float* Zn, long long ldz, cudaStream_t st);
bool project_bwd_supported(int d, int JK);
size_t project_bwd_workspace_bytes(long long n, int d, int JK);
int launch_project_bwd(const float* X, long long n, int d, long long ldx, const float* dZ, long long ldz, int JK, float* dW, void* ws,
                       size_t ws_bytes, cudaStream_t st);
int launch_rows_f32(const float* Zr, long long P, const float* Z2, long long n, long long ld, int J, int K, int base,
                    const float* c, float* out, long long ldo, cudaStream_t st);
int launch_rows_f64(const double* Zr, long long P, const double* Z2, long long n, long long ld, int J, int K, int base,
                    const double* c, double* out, long long ldo, cudaStream_t st);
int launch_reduce_dz(const float* dzp, int nsplits, long long plane, float scale, float* dz, cudaStream_t st);
int launch_reduce_g(const float* gp, long long nctas, int width, float scale, float* g, cudaStream_t st);
int launch_mvm_f64(const double* Z1, long long m, const double* Z2, long long n, long long ld, int J, int K, int base,
                   const double* c, const double* V, int t, double* out, cudaStream_t st);
int launch_quad_f64(const double* Z1, long long m, const double* Z2, long long n, long long ld, int J, int K, int base,
                    const double* c, const double* L, const double* R, int t, double* dZ1, double* g, cudaStream_t st);
int launch_axpy_rows(float alpha, const float* V, int ldv, long long m, int t, float* out, int ldo, cudaStream_t st);
int launch_reduce_partials(const float* partial, int nparts, long long m, int TP, int t, float* out, int ldo, cudaStream_t st);

}  // namespace rpgp

// aux_kernels.cu -- the small kernels around the fused K.V pair: coordinate packing, the X.W^T projection (HBM-bound,
// FP64 accumulation so the packed coordinate is a single rounding), dense rows of K, gradient reducers and the
// un-tiled FP64 path.  See include/rpgp.h for the reference interfaces each one replaces.
#include "rpgp_common.cuh"
#include <cstdlib>

#include "aux_kernels.cuh"

namespace rpgp {

// ---- packing --------------------------------------------------------------------------------------------------------
// one thread per packed element: (chunk, row, pos)
__global__ void pack_coords_kernel(const float* __restrict__ Z, long long n, long long ld, Layout lay, float scale,
                                   float* __restrict__ Zp) {
    const long long total = (long long)lay.nchunks * n * lay.CP;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int pos = (int)(idx % lay.CP);
        const long long row = (idx / lay.CP) % n;
        const int chunk = (int)(idx / ((long long)lay.CP * n));
        const int g = pos / lay.KP, mm = pos - g * lay.KP;
        const int jg = chunk * lay.G + g;
        float v = 0.f;
        if (g < lay.G && jg < lay.J && mm < lay.K) v = Z[row * ld + (long long)jg * lay.K + mm] * scale;
        Zp[idx] = v;
    }
}

__global__ void pack_log2c_kernel(const float* __restrict__ c, Layout lay, float* __restrict__ nlc) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= lay.nchunks * lay.G) return;
    nlc[idx] = (idx < lay.J) ? -log2f(c[idx]) : __int_as_float(0x7f800000);
}

// ---- projection -------------------------------------------------------------------------------------------------------
// CTA = 128 rows; W (with pre_inv folded in) is staged in shared memory as doubles [JK][d] when it fits, X rows are
// read through L1 (each row is re-used by the JK outputs of the same row).  One thread per (row, q).
__global__ void __launch_bounds__(256) project_kernel(const float* __restrict__ X, long long n, int d, long long ldx,
                                                      const float* __restrict__ W, const float* __restrict__ pre_inv,
                                                      const float* __restrict__ post_inv, Layout lay, float scale,
                                                      float* __restrict__ Zp, float* __restrict__ Zn, long long ldz, int w_in_smem) {
    extern __shared__ double wsm[];  // [JK][d] (W * pre_inv)
    const int JK = lay.J * lay.K;
    if (w_in_smem) {
        for (int e = threadIdx.x; e < JK * d; e += blockDim.x) {
            const int k = e % d;
            wsm[e] = (double)W[e] * (pre_inv ? (double)pre_inv[k] : 1.0);
        }
        __syncthreads();
    }
    const long long total = (long long)lay.nchunks * lay.CP;  // packed positions per row
    const long long row0 = (long long)blockIdx.x * 128;
    for (long long e = threadIdx.x; e < 128 * total; e += blockDim.x) {
        const long long row = row0 + e / total;
        if (row >= n) break;
        const int pp = (int)(e % total);
        const int chunk = pp / lay.CP, pos = pp - chunk * lay.CP;
        const int g = pos / lay.KP, mm = pos - g * lay.KP;
        const int jg = chunk * lay.G + g;
        float v = 0.f;
        if (g < lay.G && jg < lay.J && mm < lay.K) {
            const int q = jg * lay.K + mm;
            const float* xr = X + row * ldx;
            double acc = 0.0;
            if (w_in_smem) {
                const double* wq = wsm + (long long)q * d;
                for (int k = 0; k < d; ++k) acc = fma((double)__ldg(xr + k), wq[k], acc);
            } else {
                for (int k = 0; k < d; ++k)
                    acc = fma((double)__ldg(xr + k) * (pre_inv ? (double)pre_inv[k] : 1.0), (double)W[(long long)q * d + k], acc);
            }
            if (post_inv) acc *= (double)post_inv[q];
            v = (float)(acc * (double)scale);
            if (Zn) Zn[row * ldz + q] = (float)acc;
        }
        if (Zp) Zp[((long long)chunk * n + row) * lay.CP + pos] = v;
    }
}

// ---- base kernels on natural coordinates (sq = squared distance of a group): value f and slope df/dsq ---------------------
template <typename T>
__device__ __forceinline__ T base_f(int base, T sq) {
    if (base == 1) {
        const T q = sqrt(T(3) * sq);
        return (T(1) + q) * exp(-q);
    }
    if (base == 2) return T(1) / sqrt(sq + T(1));
    if (base == 3) return cos(sqrt(sq));
    return exp(T(-0.5) * sq);
}
template <typename T>
__device__ __forceinline__ T base_df(int base, T sq) {
    if (base == 1) return T(-1.5) * exp(-sqrt(T(3) * sq));
    if (base == 2) {
        const T r = T(1) / sqrt(sq + T(1));
        return T(-0.5) * r * r * r;
    }
    if (base == 3) {        // d cos(sqrt(sq)) / dsq = -sin(d) / (2 d)
        const T d = sqrt(sq);
        return d < T(1e-4) ? T(-0.5) + sq / T(12) : T(-0.5) * sin(d) / d;
    }
    return T(-0.5) * exp(T(-0.5) * sq);
}

// ---- dense rows of K ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void kernel_rows_kernel(const T* __restrict__ Zr, long long P, const T* __restrict__ Z2, long long n,
                                   long long ld, int J, int K, int base, const T* __restrict__ c, T* __restrict__ out,
                                   long long ldo) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long p = blockIdx.y;
    if (col >= n || p >= P) return;
    const T* a = Zr + p * ld;
    const T* b = Z2 + col * ld;
    T s = 0;
    for (int j = 0; j < J; ++j) {
        T sq = 0;
        for (int mm = 0; mm < K; ++mm) {
            const T dd = a[j * K + mm] - b[j * K + mm];
            sq += dd * dd;
        }
        s += c[j] * base_f<T>(base, sq);
    }
    out[p * ldo + col] = s;
}

// out[row, c] = sum_p partial[p, row, c]   (fixed order -> deterministic)
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nparts, long long m, int TPv, int t,
                                       float* __restrict__ out, int ldo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * t) return;
    const long long row = idx / t;
    const int c = (int)(idx - row * t);
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[((long long)p * m + row) * TPv + c];
    out[row * ldo + c] = s;
}

// out[i,c] += alpha * V[i,c]  (the sigma_n^2 I term of AddedDiagLazyTensor)
__global__ void axpy_rows_kernel(float alpha, const float* __restrict__ V, int ldv, long long m, int t, float* __restrict__ out, int ldo) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * t) return;
    const long long row = idx / t;
    const int c = (int)(idx - row * t);
    out[row * ldo + c] += alpha * V[row * ldv + c];
}

// ---- gradient reducers ------------------------------------------------------------------------------------------------
// dz[chunk][row][pos] = scale * sum_split dzp[split][chunk][row][pos]
__global__ void reduce_dz_kernel(const float* __restrict__ dzp, int nsplits, long long plane, float scale,
                                 float* __restrict__ dz) {
    // plane = nchunks*m*CP elements
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < plane;
         idx += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < nsplits; ++p) s += dzp[(long long)p * plane + idx];
        dz[idx] = s * scale;
    }
}
// g[e] = scale * sum_cta gp[cta][e],  e < nchunks*G  (double accumulation, fixed order)
__global__ void reduce_g_kernel(const float* __restrict__ gp, long long nctas, int width, float scale,
                                float* __restrict__ g) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= width) return;
    double s = 0.0;
    for (long long cta = 0; cta < nctas; ++cta) s += (double)gp[cta * width + e];
    g[e] = (float)(s * (double)scale);
}

// ---- FP64 path (un-tiled; small n) ------------------------------------------------------------------------------------
constexpr int F64_TMAX = 16;
__global__ void mvm_fwd_f64_kernel(const double* __restrict__ Z1, long long m, const double* __restrict__ Z2,
                                   long long n, long long ld, int J, int K, int base, const double* __restrict__ c,
                                   const double* __restrict__ V, int t, int t0, int tc, double* __restrict__ out) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= m) return;
    double acc[F64_TMAX];
#pragma unroll
    for (int q = 0; q < F64_TMAX; ++q) acc[q] = 0.0;
    const double* a = Z1 + row * ld;
    for (long long col = 0; col < n; ++col) {
        const double* b = Z2 + col * ld;
        double s = 0.0;
        for (int j = 0; j < J; ++j) {
            double sq = 0.0;
            for (int mm = 0; mm < K; ++mm) {
                const double dd = a[j * K + mm] - b[j * K + mm];
                sq += dd * dd;
            }
            s += c[j] * base_f<double>(base, sq);
        }
#pragma unroll
        for (int q = 0; q < F64_TMAX; ++q)
            if (q < tc) acc[q] += s * V[col * t + t0 + q];
    }
#pragma unroll
    for (int q = 0; q < F64_TMAX; ++q)
        if (q < tc) out[row * t + t0 + q] = acc[q];
}

// one thread per (row, j): accumulates dZ1[row, jK..jK+K) and atomically adds its share of g[j]
__global__ void quad_bwd_f64_kernel(const double* __restrict__ Z1, long long m, const double* __restrict__ Z2,
                                    long long n, long long ld, int J, int K, int base, const double* __restrict__ c,
                                    const double* __restrict__ L, const double* __restrict__ R, int t,
                                    double* __restrict__ dZ1, double* __restrict__ g) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * J) return;
    const long long row = idx / J;
    const int j = (int)(idx - row * J);
    const double* a = Z1 + row * ld + (long long)j * K;
    double gsum = 0.0;
    for (long long col = 0; col < n; ++col) {
        double S = 0.0;
        for (int q = 0; q < t; ++q) S += L[row * t + q] * R[col * t + q];
        const double* b = Z2 + col * ld + (long long)j * K;
        double sq = 0.0;
        for (int mm = 0; mm < K; ++mm) {
            const double dd = a[mm] - b[mm];
            sq += dd * dd;
        }
        gsum += S * base_f<double>(base, sq) * c[j];  // dG / d ln c_j
        const double w = 2.0 * S * base_df<double>(base, sq) * c[j];
        for (int mm = 0; mm < K; ++mm) dZ1[row * ld + (long long)j * K + mm] += w * (a[mm] - b[mm]);
    }
    atomicAdd(g + j, gsum);
}

// ---- launchers --------------------------------------------------------------------------------------------------------
static inline int blocks_for(long long total, int block, int cap = 148 * 32) {
    long long b = (total + block - 1) / block;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int launch_pack_coords(const float* Z, long long n, long long ld, const Layout& lay, float scale, float* Zp,
                       cudaStream_t st) {
    const long long total = (long long)lay.nchunks * n * lay.CP;
    if (total == 0) return OK;
    pack_coords_kernel<<<blocks_for(total, 256), 256, 0, st>>>(Z, n, ld, lay, scale, Zp);
    note_launch();
    return cuda_fail(cudaGetLastError(), "pack_coords_kernel");
}
int launch_pack_log2c(const float* c, const Layout& lay, float* nlc, cudaStream_t st) {
    const int total = lay.nchunks * lay.G;
    pack_log2c_kernel<<<(total + 127) / 128, 128, 0, st>>>(c, lay, nlc);
    note_launch();
    return cuda_fail(cudaGetLastError(), "pack_log2c_kernel");
}
int launch_project(const float* X, long long n, int d, long long ldx, const float* W, const float* pre_inv,
                   const float* post_inv, const Layout& lay, float scale, float* Zp, float* Zn, long long ldz, cudaStream_t st) {
    if (n == 0) return OK;
    // tensor-core path (project_tc.cu) for d <= 128, J K <= 112; RPGP_PROJECT_TC=0 keeps everything on the FP64-accumulating SIMT kernel
    static const int tc_env = [] { const char* e = getenv("RPGP_PROJECT_TC"); return e ? atoi(e) : 1; }();
    if (tc_env && project_tc_supported(d, lay)) return launch_project_tc(X, n, d, ldx, W, pre_inv, post_inv, lay, scale, Zp, Zn, ldz, st);
    const size_t wbytes = (size_t)lay.J * lay.K * d * sizeof(double);
    const int in_smem = wbytes <= 200 * 1024;
    if (in_smem && wbytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wbytes);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(project_kernel)");
    }
    const long long nblocks = (n + 127) / 128;
    project_kernel<<<(unsigned)nblocks, 256, in_smem ? wbytes : 0, st>>>(X, n, d, ldx, W, pre_inv, post_inv, lay, scale,
                                                                        Zp, Zn, ldz, in_smem);
    note_launch();
    return cuda_fail(cudaGetLastError(), "project_kernel");
}
int launch_rows_f32(const float* Zr, long long P, const float* Z2, long long n, long long ld, int J, int K, int base,
                    const float* c, float* out, long long ldo, cudaStream_t st) {
    if (P == 0 || n == 0) return OK;
    for (long long p0 = 0; p0 < P; p0 += 32768) {
        const long long pc = (P - p0 < 32768) ? (P - p0) : 32768;
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)pc);
        kernel_rows_kernel<float><<<grid, 256, 0, st>>>(Zr + p0 * ld, pc, Z2, n, ld, J, K, base, c, out + p0 * ldo, ldo);
    }
    note_launch();
    return cuda_fail(cudaGetLastError(), "kernel_rows_kernel<float>");
}
int launch_rows_f64(const double* Zr, long long P, const double* Z2, long long n, long long ld, int J, int K, int base,
                    const double* c, double* out, long long ldo, cudaStream_t st) {
    if (P == 0 || n == 0) return OK;
    for (long long p0 = 0; p0 < P; p0 += 32768) {
        const long long pc = (P - p0 < 32768) ? (P - p0) : 32768;
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)pc);
        kernel_rows_kernel<double><<<grid, 256, 0, st>>>(Zr + p0 * ld, pc, Z2, n, ld, J, K, base, c, out + p0 * ldo, ldo);
    }
    note_launch();
    return cuda_fail(cudaGetLastError(), "kernel_rows_kernel<double>");
}
int launch_reduce_partials(const float* partial, int nparts, long long m, int TP, int t, float* out, int ldo, cudaStream_t st) {
    const long long total = m * t;
    if (total == 0) return OK;
    reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, nparts, m, TP, t, out, ldo);
    note_launch();
    return cuda_fail(cudaGetLastError(), "reduce_partials_kernel");
}
int launch_axpy_rows(float alpha, const float* V, int ldv, long long m, int t, float* out, int ldo, cudaStream_t st) {
    const long long total = m * t;
    if (total == 0) return OK;
    axpy_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(alpha, V, ldv, m, t, out, ldo);
    note_launch();
    return cuda_fail(cudaGetLastError(), "axpy_rows_kernel");
}
int launch_reduce_dz(const float* dzp, int nsplits, long long plane, float scale, float* dz, cudaStream_t st) {
    if (plane == 0) return OK;
    reduce_dz_kernel<<<blocks_for(plane, 256), 256, 0, st>>>(dzp, nsplits, plane, scale, dz);
    note_launch();
    return cuda_fail(cudaGetLastError(), "reduce_dz_kernel");
}
int launch_reduce_g(const float* gp, long long nctas, int width, float scale, float* g, cudaStream_t st) {
    reduce_g_kernel<<<(width + 127) / 128, 128, 0, st>>>(gp, nctas, width, scale, g);
    note_launch();
    return cuda_fail(cudaGetLastError(), "reduce_g_kernel");
}
int launch_mvm_f64(const double* Z1, long long m, const double* Z2, long long n, long long ld, int J, int K, int base,
                   const double* c, const double* V, int t, double* out, cudaStream_t st) {
    if (m == 0) return OK;
    for (int t0 = 0; t0 < t; t0 += F64_TMAX) {
        const int tc = (t - t0 < F64_TMAX) ? (t - t0) : F64_TMAX;
        mvm_fwd_f64_kernel<<<(unsigned)((m + 63) / 64), 64, 0, st>>>(Z1, m, Z2, n, ld, J, K, base, c, V, t, t0, tc, out);
    }
    note_launch();
    return cuda_fail(cudaGetLastError(), "mvm_fwd_f64_kernel");
}
int launch_quad_f64(const double* Z1, long long m, const double* Z2, long long n, long long ld, int J, int K, int base,
                    const double* c, const double* L, const double* R, int t, double* dZ1, double* g, cudaStream_t st) {
    if (m == 0) return OK;
    const long long total = m * J;
    quad_bwd_f64_kernel<<<(unsigned)((total + 63) / 64), 64, 0, st>>>(Z1, m, Z2, n, ld, J, K, base, c, L, R, t, dZ1, g);
    note_launch();
    return cuda_fail(cudaGetLastError(), "quad_bwd_f64_kernel");
}

}  // namespace rpgp

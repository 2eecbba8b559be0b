// peaks.cu -- issue-rate microbenchmarks for the two pipes that bound the K.V path on sm_100a:
// the XU pipe (MUFU.EX2) and the FP32 FMA pipe (FFMA / FADD, scalar and packed f32x2).
//
// SURVEY.md §8(d) asks for *measured* R_xu / R_fma as roofline denominators because
// MEASURED_PEAKS.json only holds HBM and bf16-GEMM peaks.  Every kernel keeps NCH independent
// dependency chains per thread so the pipe, not latency, is the limit; all blocks are co-resident
// (one wave), and the rate is lane-ops / SM / clk with clk taken from clock64() inside the kernel.
//
// Build (standalone):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DPEAKS_MAIN peaks.cu -o peaks
// Library entry point: rpgp_measure_peaks() (declared in include/rpgp.h).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

constexpr int NCH = 8;       // independent chains per thread
constexpr int UNROLL = 8;    // inner unroll (ops per chain per outer iteration)

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ float lo32(unsigned long long v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo;
}

enum Op { FFMA = 0, FFMA2 = 1, FADD = 2, FADD2 = 3, MUFU = 4, MIX_1X_2F2 = 5, MIX_1X_4F2 = 6, MIX_1X_3F = 7, FMUL = 8 };

template <int OP>
__global__ void __launch_bounds__(256) peak_kernel(float* sink, long long* cycles, int iters, float a, float b) {
    float x[NCH];
    unsigned long long x2[NCH];
    float m[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        x[k] = a * (float)(threadIdx.x + k) * 1e-3f;
        x2[k] = pack2(x[k], x[k] + 1.0f);
        m[k] = -x[k];
    }
    const unsigned long long a2 = pack2(a, a), b2 = pack2(b, b);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if (OP == FFMA) x[k] = fmaf(x[k], a, b);
                if (OP == FMUL) x[k] = x[k] * a;
                if (OP == FFMA2) x2[k] = ffma2(x2[k], a2, b2);
                if (OP == FADD) x[k] = x[k] + a;
                if (OP == FADD2) x2[k] = fadd2(x2[k], a2);
                if (OP == MUFU) m[k] = ex2_ftz(m[k]);
                if (OP == MIX_1X_2F2) {  // 1 MUFU + 2 packed FMA per slot
                    m[k] = ex2_ftz(m[k]);
                    x2[k] = ffma2(x2[k], a2, b2);
                    x2[k] = ffma2(x2[k], a2, b2);
                }
                if (OP == MIX_1X_4F2) {  // 1 MUFU + 4 packed FMA per slot
                    m[k] = ex2_ftz(m[k]);
                    x2[k] = ffma2(x2[k], a2, b2);
                    x2[k] = ffma2(x2[k], a2, b2);
                    x2[k] = ffma2(x2[k], a2, b2);
                    x2[k] = ffma2(x2[k], a2, b2);
                }
                if (OP == MIX_1X_3F) {  // 1 MUFU + 3 scalar FP32 per slot (the un-packed K=1 inner loop mix)
                    m[k] = ex2_ftz(m[k]);
                    x[k] = fmaf(x[k], a, b);
                    x[k] = x[k] + a;
                    x[k] = fmaf(x[k], a, b);
                }
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) acc += x[k] + lo32(x2[k]) + m[k];
    if (acc == 123.456f) sink[0] = acc;  // never true; keeps the chains alive
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct OpInfo {
    const char* name;
    double fp32_lane_ops;  // FP32-pipe lane-ops per thread per (it,u,k) slot
    double mufu_ops;       // XU-pipe lane-ops per slot
};

template <int OP>
int run_one(const OpInfo& info, int sm_count, int blocks_per_sm, int iters, float* d_sink, long long* d_cycles,
            double* fp32_per_clk_sm, double* mufu_per_clk_sm, double* ms_out, double* mhz_out) {
    const int grid = sm_count * blocks_per_sm, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    peak_kernel<OP><<<grid, block>>>(d_sink, d_cycles, iters / 8, 1.0001f, 0.5f);  // warm-up
    cudaEventRecord(e0);
    peak_kernel<OP><<<grid, block>>>(d_sink, d_cycles, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) return (int)err;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(grid);
    cudaMemcpy(cyc.data(), d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long cmax = *std::max_element(cyc.begin(), cyc.end());
    const double slots_per_sm = (double)iters * UNROLL * NCH * block * blocks_per_sm;
    *fp32_per_clk_sm = info.fp32_lane_ops * slots_per_sm / (double)cmax;
    *mufu_per_clk_sm = info.mufu_ops * slots_per_sm / (double)cmax;
    *ms_out = ms;
    *mhz_out = (double)cmax / (ms * 1e3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return 0;
}

}  // namespace

// out[0..]: per op {fp32 lane-ops/clk/SM, mufu lane-ops/clk/SM, ms, effective MHz}; returns number of ops measured,
// negative on CUDA error.  names (optional) receives the op labels.
extern "C" int rpgp_measure_peaks(double* out, int max_ops, const char** names) {
    static const OpInfo infos[] = {
        {"ffma", 1, 0},          {"ffma2", 2, 0},         {"fadd", 1, 0},
        {"fadd2", 2, 0},         {"mufu_ex2", 0, 1},      {"mix_1mufu_2ffma2", 4, 1},
        {"mix_1mufu_4ffma2", 8, 1}, {"mix_1mufu_3fp32", 3, 1}, {"fmul", 1, 0},
    };
    const int nops = (int)(sizeof(infos) / sizeof(infos[0]));
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1;
    const int sms = prop.multiProcessorCount, bps = 4, iters = 2048;
    float* d_sink = nullptr;
    long long* d_cycles = nullptr;
    if (cudaMalloc(&d_sink, 4) != cudaSuccess || cudaMalloc(&d_cycles, sizeof(long long) * sms * bps) != cudaSuccess) return -2;
    int n = 0;
    for (int op = 0; op < nops && op < max_ops; ++op) {
        double* o = out + 4 * op;
        int rc = 0;
        switch (op) {
            case 0: rc = run_one<FFMA>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 1: rc = run_one<FFMA2>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 2: rc = run_one<FADD>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 3: rc = run_one<FADD2>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 4: rc = run_one<MUFU>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 5: rc = run_one<MIX_1X_2F2>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 6: rc = run_one<MIX_1X_4F2>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 7: rc = run_one<MIX_1X_3F>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
            case 8: rc = run_one<FMUL>(infos[op], sms, bps, iters, d_sink, d_cycles, o, o + 1, o + 2, o + 3); break;
        }
        if (rc != 0) { cudaFree(d_sink); cudaFree(d_cycles); return -100 - rc; }
        if (names) names[op] = infos[op].name;
        ++n;
    }
    cudaFree(d_sink);
    cudaFree(d_cycles);
    return n;
}

#ifdef PEAKS_MAIN
int main() {
    double out[4 * 16];
    const char* names[16];
    int n = rpgp_measure_peaks(out, 16, names);
    if (n < 0) { fprintf(stderr, "peaks failed: %d\n", n); return 1; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"ops\": {", prop.name, prop.multiProcessorCount);
    for (int i = 0; i < n; ++i)
        printf("%s\"%s\": {\"fp32_lane_ops_per_clk_sm\": %.2f, \"mufu_per_clk_sm\": %.2f, \"ms\": %.3f, \"mhz\": %.0f}",
               i ? ", " : "", names[i], out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
    printf("}}\n");
    return 0;
}
#endif

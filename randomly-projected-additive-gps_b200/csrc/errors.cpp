// errors.cpp -- thread-local last-error string behind rpgp_last_error() (include/rpgp.h).
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include "rpgp_common.cuh"
namespace rpgp {
static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;  // kernels launched by this library (bench.py reports it)
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void note_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
int cuda_fail(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return ERR_CUDA;
}
const char* last_error() { return g_err; }
}  // namespace rpgp

// errors.cpp -- thread-local last-error string behind rpgp_last_error() (include/rpgp.h).
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include "rpgp_common.cuh"
namespace rpgp {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return OK;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return ERR_CUDA;
}
const char* last_error() { return g_err; }
}  // namespace rpgp

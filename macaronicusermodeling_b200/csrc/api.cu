// Error reporting, version and device probe of libmlbp.so.
#include <string.h>

#include "common.cuh"

namespace mlbp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace mlbp

extern "C" const char *mlbp_last_error(void) { return mlbp::g_err; }

extern "C" int mlbp_version(void) { return 100; }

extern "C" int mlbp_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return 0;
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

/* zero n int32 words on the stream (flags and counters the kernels communicate through) */
extern "C" int mlbp_zero_words(int32_t *p, int n, void *stream) {
    if (n == 0) return MLBP_OK;
    MLBP_CHECK_ARG(p && n > 0, "zero_words: bad argument");
    MLBP_CUDA(cudaMemsetAsync(p, 0, sizeof(int32_t) * (size_t)n, mlbp::as_stream(stream)));
    return MLBP_OK;
}

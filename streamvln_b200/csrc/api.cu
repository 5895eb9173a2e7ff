// libsvln: error plumbing and device queries shared by every entry point.
#include <stdarg.h>

#include "svln_common.cuh"

namespace svln {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return SVLN_ERR_CUDA;
    }
    return SVLN_OK;
}

bool first_use_on_device(unsigned long long* mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if ((*mask >> dev) & 1ull) return false;
    *mask |= 1ull << dev;
    return true;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace svln

extern "C" {

int svln_version(void) { return 100; }  // 0.1.0

const char* svln_last_error(void) { return svln::g_err; }

int svln_sm_count(void) { return svln::sm_count(); }

}

// Library-level entry points: error string, version, device probe.
#include <string.h>

#include "common.cuh"

namespace efb {
static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
}  // namespace efb

extern "C" const char* efb_last_error(void) { return efb::g_error; }

extern "C" int efb_version(void) { return 100; /* 0.1.0 */ }

extern "C" int efb_device_count(void) {
    int n = 0;
    cudaError_t err = cudaGetDeviceCount(&n);
    if (err != cudaSuccess) {
        efb::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(err));
        cudaGetLastError();
        return -1;
    }
    return n;
}

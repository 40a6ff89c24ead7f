// Error plumbing and device queries for the C-ABI.
#include "fie_common.cuh"
#include <string.h>

namespace fie {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
static int g_sync_debug = -1;
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        if (g_sync_debug < 0) { const char* s = getenv("FIE_SYNC_DEBUG"); g_sync_debug = (s && s[0] == '1') ? 1 : 0; }
        if (g_sync_debug) e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return FIE_ERR_CUDA; }
    return FIE_OK;
}
int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}
int device_sm_count() {
    static int sms[kMaxDevices] = {0};
    const int dev = current_device();
    if (!sms[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
        sms[dev] = n;
    }
    return sms[dev];
}
}  // namespace fie

extern "C" const char* fie_last_error(void) { return fie::g_err; }
extern "C" int fie_version(void) { return 100; }
extern "C" int fie_device_supported(void) {
    int dev = 0; cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return prop.major == 10 ? 1 : 0;
}

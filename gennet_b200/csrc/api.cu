// Library-level entry points of the C ABI: error reporting, version, device check.
#include "gn_common.cuh"

namespace gn {
thread_local char g_err[512] = {0};
}

extern "C" const char* gn_last_error(void) { return gn::g_err; }

extern "C" int gn_version(void) { return 100; }

extern "C" int gn_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        snprintf(gn::g_err, sizeof(gn::g_err), "gn_device_ok: no CUDA device visible (this library has no CPU path)");
        return 0;
    }
    int dev = 0, major = 0, minor = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        snprintf(gn::g_err, sizeof(gn::g_err),
                 "gn_device_ok: device %d is sm_%d%d; the kernels are built for sm_100a (B200) only", dev, major, minor);
        return 0;
    }
    return 1;
}

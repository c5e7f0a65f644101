// Diagnostics + device probing for libmht_b200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mht {

static thread_local char g_err[512] = "";
long long g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int probe_devices() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

int check_device() {
    static int n = -1;
    if (n < 0) n = probe_devices();
    if (n <= 0) {
        set_error("no sm_100 (B200) device visible; libmht_b200 has no CPU fallback");
        return MHT_E_NODEVICE;
    }
    return MHT_OK;
}

}  // namespace mht

extern "C" int mht_version(void) { return 100; }
extern "C" const char *mht_last_error(void) { return mht::g_err; }
extern "C" int mht_device_count(void) { return mht::probe_devices(); }
extern "C" int64_t mht_launch_count(void) { return mht::g_launches; }

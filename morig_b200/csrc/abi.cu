// C-ABI plumbing: version, error string, device info.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

namespace morig {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MORIG_NO_PDL");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}
}  // namespace morig

extern "C" MORIG_API int morig_version(void) { return MORIG_ABI_VERSION; }
extern "C" MORIG_API const char *morig_last_error(void) { return morig::g_err; }
extern "C" MORIG_API int morig_sm_count(void) { return morig::sm_count(); }

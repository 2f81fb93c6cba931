// gan_track_b200 -- process-wide helpers of the C ABI: thread-local error text, cached device properties.
#include "gt_common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void gt_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int gt_num_sms() {
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

extern "C" const char* gt_last_error(void) { return g_err; }
extern "C" int gt_abi_version(void) { return 1; }
extern "C" int gt_sm_count(void) { return gt_num_sms(); }

// Tuning switch for the HBM-streaming kernels: 0 = bulk-copy staged (default), 1 = direct vector loads/stores.
static int g_stream_variant = 0;
int gt_stream_variant() { return g_stream_variant; }
extern "C" int gt_stream_config(int variant) {
    const int old = g_stream_variant;
    g_stream_variant = variant;
    return old;
}

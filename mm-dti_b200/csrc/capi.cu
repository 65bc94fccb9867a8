// Library-wide C ABI plumbing: version and thread-local error text.
#include "common.cuh"

#include <stdarg.h>
#include <stdio.h>

static thread_local char g_err[512] = "";

void mmdti_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int mmdti_version(void) { return 100; }
extern "C" const char* mmdti_last_error(void) { return g_err; }

static const unsigned long long* g_seed_off = nullptr;
const unsigned long long* mmdti_seed_offset_ptr() { return g_seed_off; }
extern "C" int mmdti_set_seed_offset(const uint64_t* device_counter) {
    g_seed_off = reinterpret_cast<const unsigned long long*>(device_counter);
    return MMDTI_OK;
}

int* mmdti_sched_slot() {
    constexpr int SLOTS = 16384;
    static int* base[MMDTI_MAX_DEVICES] = {};
    static unsigned next[MMDTI_MAX_DEVICES] = {};
    const int dev = mmdti_device_slot();
    if (!base[dev]) {
        if (cudaMalloc(&base[dev], SLOTS * 2 * sizeof(int)) != cudaSuccess) return nullptr;
        if (cudaMemset(base[dev], 0, SLOTS * 2 * sizeof(int)) != cudaSuccess) return nullptr;
    }
    return base[dev] + 2 * (next[dev]++ % SLOTS);
}

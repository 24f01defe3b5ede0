// Error reporting, version and launch accounting for the C ABI (include/b200med.h).
#include "common.cuh"
#include <stdarg.h>

namespace b200med {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
thread_local int g_sm_limit = 0;
std::atomic<int> g_pdl{1};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b200med

extern "C" __attribute__((visibility("default"))) int b200med_version(void) { return B200MED_VERSION; }
extern "C" __attribute__((visibility("default"))) const char *b200med_last_error(void) { return b200med::g_err; }
extern "C" __attribute__((visibility("default"))) int64_t b200med_launch_count(void) { return (int64_t)b200med::g_launches.load(); }
extern "C" __attribute__((visibility("default"))) int b200med_set_sm_limit(int32_t sms) {
    const int prev = b200med::g_sm_limit;
    b200med::g_sm_limit = sms > 0 ? sms : 0;
    return prev;
}
extern "C" __attribute__((visibility("default"))) int b200med_set_pdl(int32_t on) {
    return b200med::g_pdl.exchange(on ? 1 : 0);
}

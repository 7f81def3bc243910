// Error reporting / bookkeeping for the C ABI (include/pk2.h).
#include "common.cuh"

namespace pk2 {
static thread_local std::string g_err;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
}  // namespace pk2

extern "C" int pk2_version(void) { return 100; }
extern "C" const char* pk2_last_error(void) { return pk2::g_err.c_str(); }
extern "C" int64_t pk2_launch_count(void) { return (int64_t)pk2::g_launches.load(); }

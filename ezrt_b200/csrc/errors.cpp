// errors.cpp -- thread-local error string of the C ABI (the reference prints and exit(-1)s:
// P5/main.cpp:178-182, :220-225, :283-286; this library reports instead).
#include "ezrt_internal.h"

#include <cstdarg>
#include <cstdio>

static thread_local char g_error[512] = "";

extern "C" {

int ezrt_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

const char* ezrt_last_error(void) { return g_error; }

int ezrt_version(void) { return 200; }  // 200: ezrt_counters grew (node_visits ...), params.profile = 2

}  // extern "C"

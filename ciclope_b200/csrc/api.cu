// ciclope_b200 -- C-ABI plumbing: version and last-error storage.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "ciclope_b200.h"

static thread_local char g_err[512] = "";

void cfe_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void cfe_count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

// number of kernels this library has launched since it was loaded
extern "C" uint64_t cfe_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int cfe_version(void) { return CFE_ABI_VERSION; }
extern "C" const char* cfe_last_error(void) { return g_err; }

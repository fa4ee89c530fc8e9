/* Minimal stand-in for the three Win32 timing names PerfTest.cpp / Wrapper.cpp use
 * (PerfTest.cpp:155,173-184; Wrapper.cpp:67-76).  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_SHIM_WINDOWS_H
#define ORACLE_SHIM_WINDOWS_H
#include <time.h>
typedef long long __int64;
typedef union { long long QuadPart; } LARGE_INTEGER;
static inline int QueryPerformanceCounter(LARGE_INTEGER *t)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    t->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
    return 1;
}
static inline int QueryPerformanceFrequency(LARGE_INTEGER *f)
{
    f->QuadPart = 1000000000LL;
    return 1;
}
#endif

// open3dsot_b200 — library-level entry points (version, last error).
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"
#include "../../include/o3d_b200.h"

static thread_local char g_err[512] = "";

void o3d_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int o3d_version(void) { return O3D_B200_VERSION; }
extern "C" const char* o3d_last_error(void) { return g_err; }
extern "C" int o3d_opt_threads(int work) { return o3d_opt_n_threads(work); }
extern "C" int o3d_device_sms(void) { return o3d_num_sms(); }

// Error reporting, version and launch accounting of libqtx_b200.
#include <stdarg.h>

#include "common.cuh"

namespace qtx {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

}  // namespace qtx

extern "C" const char* qtx_last_error(void) { return qtx::g_err; }
extern "C" int qtx_abi_version(void) { return 1; }
extern "C" int64_t qtx_launch_count(void) { return qtx::g_launches; }
extern "C" void qtx_launch_count_reset(void) { qtx::g_launches = 0; }

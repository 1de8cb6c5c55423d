#include "common.h"

#include <cstdlib>
#include <cstring>

namespace gp {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GP_PDL");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;   // measured (profiles/r02_pdl_ab.txt): no gain inside CUDA graphs; off
  }
  return on == 1;
}

int num_sms() {
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

}  // namespace gp

extern "C" {

const char* gp_version(void) { return "gpb200 0.1 (sm_100a)"; }
const char* gp_last_error(void) { return gp::g_err; }
uint64_t gp_launch_count(void) { return gp::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"

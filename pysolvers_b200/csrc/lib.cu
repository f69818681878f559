// Library-wide state: error text, launch counter, cached device properties.
#include "common.cuh"

#include <stdarg.h>

namespace psb {

static thread_local char t_error[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

struct DevProps { int device = -1; int sms = 0; int smem_optin = 0; };
static thread_local DevProps t_props;

static void refresh_props() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  if (dev == t_props.device) return;
  t_props.device = dev;
  cudaDeviceGetAttribute(&t_props.sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&t_props.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
}

int sm_count() {
  refresh_props();
  return t_props.sms > 0 ? t_props.sms : 148;
}

int max_optin_smem() {
  refresh_props();
  return t_props.smem_optin > 0 ? t_props.smem_optin : 227 * 1024;
}

}  // namespace psb

extern "C" int psb_version(void) { return 100; }

extern "C" const char* psb_last_error(void) { return psb::t_error; }

extern "C" long long psb_launch_count(void) {
  return psb::g_launches.load(std::memory_order_relaxed);
}

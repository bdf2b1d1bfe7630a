// Error plumbing and device queries shared by every translation unit of libspacer_b200.so.
#include "common.cuh"
#include "spacer_b200.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include <atomic>

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};
}

void sb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sb_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    sb_set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return 1;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

namespace {
int g_pdl = -1;
}
bool sb_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("SB_NO_PDL");
    g_pdl = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return g_pdl == 1;
}
extern "C" int sb_set_pdl(int enable) {
  g_pdl = enable ? 1 : 0;
  return 0;
}

int sb_trace_set_gemm(unsigned long long*, int);
int sb_trace_set_decode(unsigned long long*, int);
int sb_trace_set_dec_attn(unsigned long long*, int);
int sb_trace_set_sampler(unsigned long long*, int);

extern "C" int sb_trace_enable(unsigned long long* buf_dev, int capacity) {
  SB_REQUIRE((buf_dev == nullptr) == (capacity == 0), "sb_trace_enable: pass (NULL, 0) to disable");
  SB_REQUIRE(!sb_trace_set_gemm(buf_dev, capacity) && !sb_trace_set_decode(buf_dev, capacity) &&
                 !sb_trace_set_dec_attn(buf_dev, capacity) && !sb_trace_set_sampler(buf_dev, capacity),
             "sb_trace_enable: cudaMemcpyToSymbol failed");
  return 0;
}

extern "C" int sb_launch_counter(long long* count_host, int reset) {
  if (count_host) *count_host = g_launches.load(std::memory_order_relaxed);
  if (reset) g_launches.store(0, std::memory_order_relaxed);
  return 0;
}

extern "C" const char* sb_last_error(void) { return g_err; }
extern "C" int sb_abi_version(void) { return SB_ABI_VERSION; }

extern "C" int sb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, mnr = 0;
  SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SB_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  SB_CUDA(cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = mnr;
  SB_REQUIRE(maj == 10, "spacer_b200 kernels are built for sm_100a only; device is sm_%d%d", maj, mnr);
  return 0;
}

// Library bookkeeping: version, launch counter, CUDA error capture, host<->device staging.
#include "ifx_host.cuh"
#include <atomic>
#include <stdio.h>

static std::atomic<int64_t> g_launches{0};
static char g_err[256] = "";
static std::mutex g_scratch_mu;
static void* g_buf[8] = {nullptr};
static size_t g_cap[8] = {0};

void ifx_count_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#ifdef IFX_HOSTSIM
int ifx_set_cuda_error(int err) { return err ? IFX_E_CUDA : IFX_OK; }
int ifx_after_launch() { ifx_count_launches(1); return IFX_OK; }
int ifx_h2d(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
int ifx_d2h(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); return 0; }
int ifx_sync() { return IFX_OK; }
static void* dev_alloc(size_t b) { return malloc(b); }
static void dev_free(void* p) { free(p); }
#else
int ifx_set_cuda_error(int err) {
  if (err == 0) return IFX_OK;
  snprintf(g_err, sizeof(g_err), "%s", cudaGetErrorString((cudaError_t)err));
  return IFX_E_CUDA;
}
int ifx_after_launch() {
  ifx_count_launches(1);
  return ifx_set_cuda_error((int)cudaGetLastError());
}
int ifx_h2d(void* dst, const void* src, size_t bytes) {
  return ifx_set_cuda_error((int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, nullptr));
}
int ifx_d2h(void* dst, const void* src, size_t bytes) {
  return ifx_set_cuda_error((int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, nullptr));
}
int ifx_sync() { return ifx_set_cuda_error((int)cudaStreamSynchronize(nullptr)); }
static void* dev_alloc(size_t b) {
  void* p = nullptr;
  if (ifx_set_cuda_error((int)cudaMalloc(&p, b))) return nullptr;
  return p;
}
static void dev_free(void* p) { cudaFree(p); }
#endif

IfxScratch::IfxScratch() { g_scratch_mu.lock(); }
IfxScratch::~IfxScratch() { g_scratch_mu.unlock(); }
void* IfxScratch::take(size_t bytes) {
  if (next_ >= 8) { ok_ = false; return nullptr; }
  const int i = next_++;
  if (g_cap[i] < bytes) {
    if (g_buf[i]) dev_free(g_buf[i]);
    size_t cap = bytes + bytes / 2;
    cap = (cap + 4095) & ~(size_t)4095;
    g_buf[i] = dev_alloc(cap);
    g_cap[i] = g_buf[i] ? cap : 0;
    if (!g_buf[i]) ok_ = false;
  }
  return g_buf[i];
}

extern "C" {
int ifx_version(void) { return 100; }
const char* ifx_last_cuda_error(void) { return g_err; }
int64_t ifx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
int ifx_is_hostsim(void) {
#ifdef IFX_HOSTSIM
  return 1;
#else
  return 0;
#endif
}
}

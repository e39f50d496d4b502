// Host-side helpers shared by the C ABI translation units: launch bookkeeping, error capture and the
// grow-only device scratch used by the *_host entry points.
#pragma once
#include "ifx_common.cuh"
#include <stdlib.h>
#include <mutex>

int ifx_after_launch();                 // counts the launch, captures launch errors -> IFX_OK / IFX_E_CUDA
void ifx_count_launches(int64_t n);
int ifx_set_cuda_error(int err);        // remembers cudaGetErrorString; returns IFX_E_CUDA (or IFX_OK for 0)
int ifx_h2d(void* dst, const void* src, size_t bytes);
int ifx_d2h(void* dst, const void* src, size_t bytes);
int ifx_sync();

// Device scratch for the host-buffer entry points.  Up to 8 buffers, each grow-only and reused across
// calls; the lock makes the *_host calls safe to use from several Python threads.
struct IfxScratch {
  IfxScratch();
  ~IfxScratch();
  void* take(size_t bytes);
  bool ok() const { return ok_; }
 private:
  int next_ = 0;
  bool ok_ = true;
};

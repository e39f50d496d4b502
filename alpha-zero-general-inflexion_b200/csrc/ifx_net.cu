// Input stage of the policy/value net: im2col of the 4 x 7 x 7 planes for conv1 (3x3, pad 1; InflexionNNet.py:24,41),
// written as binary16 rows [n*49][64] (36 taps c*9 + ky*3 + kx, zero-padded to 64 = one 128-byte row), so that conv1
// becomes one tensor-core GEMM [n*49, 64] x [64, 512] with a fused bias+ReLU epilogue whose output IS the NHWC
// activation conv2 reads.  cuDNN's own kernel for a 4-channel fp16 convolution is a legacy non-tensor-core path that
// took 12 % of the self-play step for 0.45 % of the net's FLOPs (profiles/r1_launches_bench_default.txt).
// One thread per (row, 8-tap chunk): a 16-byte store each, reads served by L1/L2 (784 B of planes per position).
#include "ifx_host.cuh"

#ifndef IFX_HOSTSIM
#include <cuda_fp16.h>
__global__ void __launch_bounds__(256) net_im2col3x3_kernel(const float* __restrict__ planes, int64_t rows, uint4* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * 8) return;
  const int64_t row = t >> 3;
  const int chunk = (int)(t & 7);
  const int64_t b = row / IFX_CELLS;
  const int p = (int)(row - b * IFX_CELLS), y = p / IFX_N, x = p - y * IFX_N;
  const float* src = planes + b * (4 * IFX_CELLS);
  __half h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = chunk * 8 + j;
    float v = 0.0f;
    if (k < 36) {
      const int c = k / 9, r = k - c * 9, yy = y + r / 3 - 1, xx = x + (r - (r / 3) * 3) - 1;
      if (yy >= 0 && yy < IFX_N && xx >= 0 && xx < IFX_N) v = src[c * IFX_CELLS + yy * IFX_N + xx];
    }
    h[j] = __float2half_rn(v);                               // 0/1 masks and the turn counter (<= 343) are exact
  }
  out[t] = *reinterpret_cast<const uint4*>(h);
}
#endif

extern "C" int ifx_net_im2col3x3(const float* planes, int64_t n, void* out, void* stream) {
  if (n < 0 || (n > 0 && (!planes || !out))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  (void)stream;
  return IFX_E_NODEVICE;                                     // binary16 tensor-core input stage: GPU only
#else
  const int64_t threads = n * IFX_CELLS * 8;
  net_im2col3x3_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(planes, n * IFX_CELLS,
                                                                                            reinterpret_cast<uint4*>(out));
  return ifx_after_launch();
#endif
}

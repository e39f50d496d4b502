// Input stage of the policy/value net: im2col of the 4 x 7 x 7 planes for conv1 (3x3, pad 1; InflexionNNet.py:24,41),
// written as binary16 rows [n*49][48] (36 taps c*9 + ky*3 + kx, zero-padded to 48 = three 32-byte sectors), so that
// conv1 becomes one tensor-core GEMM [n*49, 48] x [48, 512] with a fused bias+ReLU epilogue whose output IS the NHWC
// activation conv2 reads.  cuDNN's own kernel for a 4-channel fp16 convolution is a legacy non-tensor-core path that
// took 12 % of the self-play step for 0.45 % of the net's FLOPs (profiles/r1_launches_bench_default.txt).
//
// One thread per board position: the planes of the block's five games are staged in shared memory once, each thread
// gathers its 3x3 window of the four planes from there (taps outside the board read a zero slot) and writes its
// 96-byte row as six 16-byte stores; consecutive threads write consecutive rows, so a warp covers 3 KB of contiguous
// memory.  The 19 MB a batch of 4 096 games writes stay in L2 (ncu: no DRAM writes) for the GEMM that follows.
// Measured on B200 for 4 096 games: 9.5 us (round 1's form -- one thread per 8-tap chunk with index arithmetic per
// tap, rows padded to 64 -- took 18.2 us, more than the tree wave; staging the rows in a shared-memory tile for
// fully coalesced stores was slower, 12 us: the kernel is bound by its gather instructions, not by the stores).
#include "ifx_host.cuh"

#define IM2COL_GAMES 5                       // games per block: 5 x 49 = 245 of 256 threads busy
#define IM2COL_K 48                          // halfs per output row

#ifndef IFX_HOSTSIM
#include <cuda_fp16.h>
__global__ void __launch_bounds__(256) net_im2col3x3_kernel(const float* __restrict__ planes, int64_t n, uint4* __restrict__ out) {
  __shared__ float sh[IM2COL_GAMES][4 * IFX_CELLS + 4];       // + a zero slot at index 196 for taps outside the board
  const int64_t b0 = (int64_t)blockIdx.x * IM2COL_GAMES;
  const int games = (int)min((int64_t)IM2COL_GAMES, n - b0);
  for (int i = threadIdx.x; i < games * 4 * IFX_CELLS; i += 256) {
    const int gi = i / (4 * IFX_CELLS);
    sh[gi][i - gi * 4 * IFX_CELLS] = planes[b0 * (4 * IFX_CELLS) + i];
  }
  if (threadIdx.x < IM2COL_GAMES) sh[threadIdx.x][4 * IFX_CELLS] = 0.0f;
  __syncthreads();
  const int gi = threadIdx.x / IFX_CELLS, p = threadIdx.x - gi * IFX_CELLS;
  if (gi >= games) return;
  const int y = p / IFX_N, x = p - y * IFX_N;
  int off[9];                                                 // offset of the 9 window cells inside a plane, -1 = outside
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const int yy = y + r / 3 - 1, xx = x + r % 3 - 1;
    off[r] = (yy >= 0 && yy < IFX_N && xx >= 0 && xx < IFX_N) ? yy * IFX_N + xx : -1;
  }
  const float* s = sh[gi];
  __half2 h[IM2COL_K / 2];
#pragma unroll
  for (int k = 0; k < 36; k += 2) {                           // column k = c*9 + r
    const int c0 = k / 9, r0 = k - c0 * 9, c1 = (k + 1) / 9, r1 = (k + 1) - c1 * 9;
    const float v0 = s[off[r0] < 0 ? 4 * IFX_CELLS : c0 * IFX_CELLS + off[r0]];
    const float v1 = s[off[r1] < 0 ? 4 * IFX_CELLS : c1 * IFX_CELLS + off[r1]];
    h[k / 2] = __floats2half2_rn(v0, v1);                     // 0/1 masks and the turn counter (<= 343) are exact
  }
#pragma unroll
  for (int k = 18; k < IM2COL_K / 2; ++k) h[k] = __floats2half2_rn(0.0f, 0.0f);
  uint4* dst = out + ((b0 + gi) * IFX_CELLS + p) * (IM2COL_K * 2 / 16);
  const uint4* src = reinterpret_cast<const uint4*>(h);
#pragma unroll
  for (int q = 0; q < IM2COL_K * 2 / 16; ++q) dst[q] = src[q];
}
#endif

extern "C" int ifx_net_im2col3x3(const float* planes, int64_t n, void* out, void* stream) {
  if (n < 0 || (n > 0 && (!planes || !out)) || ((uintptr_t)out & 15)) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  (void)stream;
  return IFX_E_NODEVICE;                                     // binary16 tensor-core input stage: GPU only
#else
  const int64_t blocks = (n + IM2COL_GAMES - 1) / IM2COL_GAMES;
  net_im2col3x3_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(planes, n, reinterpret_cast<uint4*>(out));
  return ifx_after_launch();
#endif
}

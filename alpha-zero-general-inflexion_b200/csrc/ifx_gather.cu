// Leaf gather: packs the rows of the games that are still searching into one contiguous batch for the policy/value net,
// and scatters the net's answers back (BASELINE.json north_star, kernel 3; the reference evaluates one leaf per real
// simulation only, MCTS.py:89-93 -- a finished game never reaches the net).
//
// Under continuous self-play every game has a leaf pending in >= 99 % of the waves, so the tree kernel writes game g's
// planes to row g and the batch is dense as it is.  At the tail of a fixed set of episodes, and in Arena (where only
// the games whose side to move is the MCTS player are searched), most rows are idle: `ifx_mcts_live_rows` lists the
// games that will search in the coming move (ascending, deterministic), and these two kernels move only those rows, so
// the net runs on ceil(live / q) * q rows instead of G.
//
// One thread per 16 bytes of a row (4 bytes when a row, stride or base is not 16-byte aligned: the value vector).  The row count is read from DEVICE memory (n_ptr), so a CUDA graph
// captured for a bucket of `rows_cap` rows serves every move whose live count fits the bucket.
#include "ifx_host.cuh"

#ifndef IFX_HOSTSIM
template <class V> __device__ __forceinline__ V gather_zero();
template <> __device__ __forceinline__ IfxVec16 gather_zero<IfxVec16>() { IfxVec16 z = {0u, 0u, 0u, 0u}; return z; }
template <> __device__ __forceinline__ uint32_t gather_zero<uint32_t>() { return 0u; }

// pack: dst[k] = src[idx[k]] for k < *n_ptr, zero rows for *n_ptr <= k < rows_cap (the net sees defined input)
template <class V>
__global__ void __launch_bounds__(256) gather_pack_kernel(const char* __restrict__ src, int64_t src_stride, const int32_t* __restrict__ idx,
                                                          const int32_t* __restrict__ n_ptr, int64_t rows_cap, int chunks,
                                                          char* __restrict__ dst, int64_t dst_stride) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows_cap * chunks) return;
  const int64_t k = t / chunks;
  const int c = (int)(t - k * chunks);
  V v = gather_zero<V>();
  if (k < *n_ptr) v = *reinterpret_cast<const V*>(src + (int64_t)idx[k] * src_stride + sizeof(V) * c);
  *reinterpret_cast<V*>(dst + k * dst_stride + sizeof(V) * c) = v;
}
// scatter: dst[idx[k]] = src[k] for k < *n_ptr
template <class V>
__global__ void __launch_bounds__(256) gather_scatter_kernel(const char* __restrict__ src, int64_t src_stride, const int32_t* __restrict__ idx,
                                                             const int32_t* __restrict__ n_ptr, int64_t rows_cap, int chunks,
                                                             char* __restrict__ dst, int64_t dst_stride) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows_cap * chunks) return;
  const int64_t k = t / chunks;
  const int c = (int)(t - k * chunks);
  if (k < *n_ptr)
    *reinterpret_cast<V*>(dst + (int64_t)idx[k] * dst_stride + sizeof(V) * c) = *reinterpret_cast<const V*>(src + k * src_stride + sizeof(V) * c);
}
#endif

// element size the copy runs at: 16 bytes when rows, strides and bases allow it, else 4; 0 = bad arguments
static int gather_unit(const void* src, const void* dst, const int32_t* idx, const int32_t* n_ptr, int64_t rows_cap,
                       int64_t row_bytes, int64_t src_stride, int64_t dst_stride) {
  if (!src || !dst || !idx || !n_ptr || rows_cap < 0 || row_bytes <= 0) return 0;
  if (src_stride < row_bytes || dst_stride < row_bytes) return 0;
  const uintptr_t all = (uintptr_t)row_bytes | (uintptr_t)src_stride | (uintptr_t)dst_stride | (uintptr_t)src | (uintptr_t)dst;
  if (all & 3) return 0;
  return (all & 15) ? 4 : 16;
}

extern "C" int ifx_gather_pack_rows(const void* src, int64_t src_stride, const int32_t* idx, const int32_t* n_ptr,
                                    int64_t rows_cap, int64_t row_bytes, void* dst, int64_t dst_stride, void* stream) {
  const int unit = gather_unit(src, dst, idx, n_ptr, rows_cap, row_bytes, src_stride, dst_stride);
  if (!unit) return IFX_E_ARG;
  if (rows_cap == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  (void)stream;
  for (int64_t k = 0; k < rows_cap; ++k) {
    if (k < *n_ptr) memcpy((char*)dst + k * dst_stride, (const char*)src + (int64_t)idx[k] * src_stride, (size_t)row_bytes);
    else memset((char*)dst + k * dst_stride, 0, (size_t)row_bytes);
  }
  return ifx_after_launch();
#else
  const int chunks = (int)(row_bytes / unit);
  const int64_t threads = rows_cap * chunks;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (unit == 16)
    gather_pack_kernel<IfxVec16><<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src, src_stride, idx, n_ptr, rows_cap, chunks, (char*)dst, dst_stride);
  else
    gather_pack_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src, src_stride, idx, n_ptr, rows_cap, chunks, (char*)dst, dst_stride);
  return ifx_after_launch();
#endif
}

extern "C" int ifx_gather_scatter_rows(const void* src, int64_t src_stride, const int32_t* idx, const int32_t* n_ptr,
                                       int64_t rows_cap, int64_t row_bytes, void* dst, int64_t dst_stride, void* stream) {
  const int unit = gather_unit(src, dst, idx, n_ptr, rows_cap, row_bytes, src_stride, dst_stride);
  if (!unit) return IFX_E_ARG;
  if (rows_cap == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  (void)stream;
  for (int64_t k = 0; k < rows_cap && k < *n_ptr; ++k)
    memcpy((char*)dst + (int64_t)idx[k] * dst_stride, (const char*)src + k * src_stride, (size_t)row_bytes);
  return ifx_after_launch();
#else
  const int chunks = (int)(row_bytes / unit);
  const int64_t threads = rows_cap * chunks;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (unit == 16)
    gather_scatter_kernel<IfxVec16><<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src, src_stride, idx, n_ptr, rows_cap, chunks, (char*)dst, dst_stride);
  else
    gather_scatter_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const char*)src, src_stride, idx, n_ptr, rows_cap, chunks, (char*)dst, dst_stride);
  return ifx_after_launch();
#endif
}

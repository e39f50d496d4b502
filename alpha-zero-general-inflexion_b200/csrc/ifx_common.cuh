// Common definitions for the inflexion_b200 CUDA library (sm_100a).
//
// The same sources also compile as plain C++ with -DIFX_HOSTSIM.  That build exists ONLY so the kernel
// logic can be checked in a container without a GPU (tests/hostsim); it emulates each warp as 32
// sequential lanes.  The product never loads it: the Python package loads libifx_b200.so (nvcc, sm_100a)
// and fails loudly if it or a CUDA device is missing.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#include "../../include/ifx_b200.h"

#ifdef IFX_HOSTSIM
  #define IFX_HD static inline
  #define IFX_D static inline
  #define IFX_D_NOINLINE static
  #define IFX_LANE0 true
  #define IFX_FOR_LANES(lane) for (int lane = 0; lane < 32; ++lane)
  #define IFX_SYNCWARP() do {} while (0)
  static inline int ifx_popc64(uint64_t x) { return __builtin_popcountll(x); }
  static inline int ifx_popc32(uint32_t x) { return __builtin_popcount(x); }
  static inline int ifx_ffs32(uint32_t x) { return __builtin_ffs((int)x); }
  // host build is compiled with -ffp-contract=off, so these are single IEEE operations
  static inline float ifx_fmul(float a, float b) { return a * b; }
  static inline float ifx_fadd(float a, float b) { return a + b; }
  static inline float ifx_fdiv(float a, float b) { return a / b; }
  static inline double ifx_dmul(double a, double b) { return a * b; }
  static inline double ifx_dadd(double a, double b) { return a + b; }
  static inline double ifx_ddiv(double a, double b) { return a / b; }
  static inline double ifx_dsqrt(double a) { return sqrt(a); }
  static inline float ifx_fsqrt(float a) { return sqrtf(a); }
#else
  #include <cuda_runtime.h>
  #define IFX_HD __device__ __forceinline__
  #define IFX_D __device__ __forceinline__
  #define IFX_D_NOINLINE __device__ __noinline__   // rare paths kept out of the hot kernel's register allocation
  #define IFX_LANE0 ((threadIdx.x & 31u) == 0u)
  #define IFX_FOR_LANES(lane) for (int lane = (int)(threadIdx.x & 31u), _ifx_once = 1; _ifx_once; _ifx_once = 0)
  #define IFX_SYNCWARP() __syncwarp()
  IFX_D int ifx_popc64(uint64_t x) { return __popcll(x); }
  IFX_D int ifx_popc32(uint32_t x) { return __popc(x); }
  IFX_D int ifx_ffs32(uint32_t x) { return __ffs((int)x); }
  // explicit round-to-nearest intrinsics: never contracted into FMA (numpy does separate operations)
  IFX_D float ifx_fmul(float a, float b) { return __fmul_rn(a, b); }
  IFX_D float ifx_fadd(float a, float b) { return __fadd_rn(a, b); }
  IFX_D float ifx_fdiv(float a, float b) { return __fdiv_rn(a, b); }
  IFX_D double ifx_dmul(double a, double b) { return __dmul_rn(a, b); }
  IFX_D double ifx_dadd(double a, double b) { return __dadd_rn(a, b); }
  IFX_D double ifx_ddiv(double a, double b) { return __ddiv_rn(a, b); }
  IFX_D double ifx_dsqrt(double a) { return __dsqrt_rn(a); }
  IFX_D float ifx_fsqrt(float a) { return __fsqrt_rn(a); }
#endif

#ifdef IFX_HOSTSIM
static inline uint32_t ifx_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float ifx_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#else
IFX_D uint32_t ifx_f2u(float f) { return __float_as_uint(f); }
IFX_D float ifx_u2f(uint32_t u) { return __uint_as_float(u); }
#endif

// IEEE binary16 -> float32 (the net's head output in the fused-softmax mode)
#ifdef IFX_HOSTSIM
struct uint2 { uint32_t x, y; };
static inline float ifx_h2f(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else { int e = -1; uint32_t m = man; do { ++e; m <<= 1; } while (!(m & 0x400u)); bits = sign | ((uint32_t)(112 - e) << 23) | ((m & 0x3FFu) << 13); }
  } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
  else bits = sign | ((exp + 112) << 23) | (man << 13);
  float f; memcpy(&f, &bits, 4); return f;
}
static inline float ifx_expf(float x) { return expf(x); }
static inline float ifx_fast_exp2f(float x) { return exp2f(x); }
static inline float ifx_frcp(float x) { return 1.0f / x; }
static inline float ifx_logf(float x) { return logf(x); }
static inline float ifx_tanhf(float x) { return tanhf(x); }
static inline float ifx_tanh_fast(float x) { return tanhf(x); }
static inline float ifx_cosf(float x) { return cosf(x); }
static inline float ifx_powf(float x, float y) { return powf(x, y); }
static inline float ifx_noise_logf(float x) { return logf(x); }
#else
#include <cuda_fp16.h>
IFX_D float ifx_h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
IFX_D float ifx_expf(float x) { return expf(x); }
IFX_D float ifx_fast_exp2f(float x) {                           // one MUFU.EX2 (2 ulp, flush-to-zero); fused-head path only
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
IFX_D float ifx_frcp(float x) { return __frcp_rn(x); }
IFX_D float ifx_logf(float x) { return logf(x); }
IFX_D float ifx_tanhf(float x) { return tanhf(x); }
// tanh(x) = 1 - 2 / (2^(2x log2 e) + 1) with MUFU.EX2 / MUFU.RCP: absolute error ~2e-7 (fused-head path only, whose
// agreement with the fp32 module is a 1e-3 tolerance); saturates correctly at +-1
IFX_D float ifx_tanh_fast(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}
// noise-only helpers: the fast intrinsics (no slow paths, no local memory); their accuracy is ample for a random draw
IFX_D float ifx_cosf(float x) { return __cosf(x); }
IFX_D float ifx_powf(float x, float y) { return __powf(x, y); }
IFX_D float ifx_noise_logf(float x) { return __logf(x); }
#endif

// ---------------------------------------------------------------------------------------------
// Per-lane variables and warp collectives.  On the device a LaneVar is one register per thread; in
// the host simulation it is an array of 32 and the collectives loop.
// ---------------------------------------------------------------------------------------------
template <class T> struct LaneVar {
#ifdef IFX_HOSTSIM
  T v[32];
  T& operator[](int l) { return v[l]; }
  const T& operator[](int l) const { return v[l]; }
#else
  T v;
  IFX_D T& operator[](int) { return v; }
  IFX_D const T& operator[](int) const { return v; }
#endif
};

#ifdef IFX_HOSTSIM
static inline uint32_t lanes_ballot(const LaneVar<int>& p) {
  uint32_t m = 0;
  for (int l = 0; l < 32; ++l) if (p[l]) m |= 1u << l;
  return m;
}
static inline int lanes_sum_i32(const LaneVar<int>& x) {
  int s = 0;
  for (int l = 0; l < 32; ++l) s += x[l];
  return s;
}
static inline int lanes_max_i32(const LaneVar<int>& x) {
  int s = x[0];
  for (int l = 1; l < 32; ++l) s = x[l] > s ? x[l] : s;
  return s;
}
// x[l] <- x[l] + x[l ^ mask]   (simultaneous butterfly step; float add is commutative so both partners agree)
static inline void lanes_bfly_fadd(LaneVar<float>& x, int mask) {
  float t[32];
  for (int l = 0; l < 32; ++l) t[l] = x[l] + x[l ^ mask];
  for (int l = 0; l < 32; ++l) x[l] = t[l];
}
template <class T> static inline T lanes_bcast(const LaneVar<T>& x, int src) { return x[src]; }
static inline float lanes_max_f32(const LaneVar<float>& x) {
  float m = x[0];
  for (int l = 1; l < 32; ++l) m = x[l] > m ? x[l] : m;
  return m;
}
// butterfly sum (same association order as the device shuffles)
static inline float lanes_sum_f32(const LaneVar<float>& x) {
  float t[32];
  for (int l = 0; l < 32; ++l) t[l] = x[l];
  for (int off = 16; off > 0; off >>= 1) { float u[32]; for (int l = 0; l < 32; ++l) u[l] = t[l] + t[l ^ off]; for (int l = 0; l < 32; ++l) t[l] = u[l]; }
  return t[0];
}
// out[l] = in[l + 1]  (lane 31 keeps its own value)
static inline void lanes_shift_down(const LaneVar<uint32_t>& in, LaneVar<uint32_t>& out) {
  for (int l = 0; l < 32; ++l) out[l] = in[l < 31 ? l + 1 : 31];
}
// maximum u; ties go to the LOWEST action index (sequential scan with strict '>' in MCTS.py:119-129)
static inline void lanes_argmax(const LaneVar<float>& u, const LaneVar<int>& a, float* bu, int* ba) {
  float best = u[0]; int arg = a[0];
  for (int l = 1; l < 32; ++l)
    if (u[l] > best || (u[l] == best && a[l] < arg)) { best = u[l]; arg = a[l]; }
  *bu = best; *ba = arg;
}
#else
IFX_D uint32_t lanes_ballot(const LaneVar<int>& p) { return __ballot_sync(0xffffffffu, p.v != 0); }
IFX_D int lanes_sum_i32(const LaneVar<int>& x) { return __reduce_add_sync(0xffffffffu, x.v); }
IFX_D int lanes_max_i32(const LaneVar<int>& x) { return __reduce_max_sync(0xffffffffu, x.v); }
IFX_D void lanes_bfly_fadd(LaneVar<float>& x, int mask) {
  x.v = __fadd_rn(x.v, __shfl_xor_sync(0xffffffffu, x.v, mask));
}
template <class T> IFX_D T lanes_bcast(const LaneVar<T>& x, int src) { return __shfl_sync(0xffffffffu, x.v, src); }
IFX_D float lanes_max_f32(const LaneVar<float>& x) {          // one REDUX on keys that sort like the floats (NaN-free input)
  uint32_t k = __float_as_uint(x.v);
  k = (k & 0x80000000u) ? ~k : (k | 0x80000000u);
  const uint32_t kmax = __reduce_max_sync(0xffffffffu, k);
  return __uint_as_float((kmax & 0x80000000u) ? (kmax & 0x7fffffffu) : ~kmax);
}
IFX_D float lanes_sum_f32(const LaneVar<float>& x) {
  float t = x.v;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) t = __fadd_rn(t, __shfl_xor_sync(0xffffffffu, t, off));
  return t;
}
IFX_D void lanes_shift_down(const LaneVar<uint32_t>& in, LaneVar<uint32_t>& out) {
  out.v = __shfl_down_sync(0xffffffffu, in.v, 1);
}
// Two REDUX operations instead of a five-step shuffle butterfly: floats map to unsigned keys that sort the
// same way (with -0.0 folded onto +0.0, which compare equal), the maximum key is found with redux.max, and the
// lowest action index among the lanes holding it with redux.min.
IFX_D void lanes_argmax(const LaneVar<float>& u, const LaneVar<int>& a, float* bu, int* ba) {
  uint32_t k = __float_as_uint(u.v);
  k = (k == 0x80000000u) ? 0u : k;
  k = (k & 0x80000000u) ? ~k : (k | 0x80000000u);
  const uint32_t kmax = __reduce_max_sync(0xffffffffu, k);
  const uint32_t cand = (k == kmax) ? (uint32_t)a.v : 0xffffffffu;
  const uint32_t amin = __reduce_min_sync(0xffffffffu, cand);
  const uint32_t back = (kmax & 0x80000000u) ? (kmax & 0x7fffffffu) : ~kmax;
  *bu = __uint_as_float(back); *ba = (int)amin;
}
#endif

// ---------------------------------------------------------------------------------------------
// constants of the game (inflexion/InflexionGame.py:14-21, 47-50, 68; main.py:34)
// ---------------------------------------------------------------------------------------------
struct alignas(16) IfxVec16 { uint32_t x, y, z, w; };

#define IFX_N 7
#define IFX_CELLS 49
#define IFX_MOVES 7
#define IFX_ACTIONS 343
#define IFX_SPAWN 6
#define IFX_MAX_POWER_AT_SPAWN 48
#define IFX_BOARD_MASK ((1ull << 49) - 1ull)
#define IFX_ROW 352            // padded action row (343 -> 352 floats = 11 x 128 B)

// Deterministic policy/value stub as ONE kernel (stubnet.py:stub_policy_value, bit for bit): the evaluator of the
// bit-exact parity runs ("a fixed deterministic net stub", BASELINE.json north_star) and of the tree-and-rules-only
// throughput runs, where a torch implementation (a dozen elementwise launches per wave) would dominate the step.
// Integer mixing, one int->float32 conversion and one IEEE float32 division per prior: nothing whose rounding is
// implementation defined.  One warp per game.
#include "ifx_host.cuh"

#define STUB_MASK31 0x7FFFFFFFll

IFX_HD int64_t stub_mix(int64_t h) {
  h = (h ^ (h >> 15)) & STUB_MASK31;
  h = (h * 0x2C1B3C6Dll) & STUB_MASK31;
  h = (h ^ (h >> 12)) & STUB_MASK31;
  h = (h * 0x297A2D39ll) & STUB_MASK31;
  return (h ^ (h >> 15)) & STUB_MASK31;
}

IFX_D void body_stub(const float* x, const int64_t* coef, int dyadic, float* pi, float* v) {
  LaneVar<int64_t> part;
  IFX_FOR_LANES(l) {
    int64_t s = 0;
    for (int i = l; i < 196; i += 32) s += (int64_t)x[i] * coef[i];        // planes are small integers held in float32
    part[l] = s;
  }
#ifdef IFX_HOSTSIM
  int64_t tot = 0;
  for (int l = 0; l < 32; ++l) tot += part[l];
#else
  int64_t tot = part.v;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
#endif
  const int64_t h = stub_mix(tot & STUB_MASK31);
  IFX_FOR_LANES(l) {
    for (int a = l; a < IFX_ACTIONS; a += 32) {
      const int64_t z = stub_mix((h + (int64_t)a * 0x9E3779Bll) & STUB_MASK31);
      const int64_t sharp = (z & 0xFF) + 1;                                   // 1..256
      if (dyadic) {
        pi[a] = (float)sharp;
      } else {
        const int64_t peaked = ((z >> 8) % 11 == 0) ? sharp * 64 : sharp;
        pi[a] = ifx_fdiv((float)peaked, 1000003.0f);
      }
    }
  }
  if (IFX_LANE0) {
    const int64_t hv = stub_mix(h ^ 0x5BD1E995ll);
    *v = ifx_fdiv((float)((hv & 0xFFFF) - 32768), 32768.0f);                  // dyadic in [-1, 1)
  }
}

#ifndef IFX_HOSTSIM
__global__ void __launch_bounds__(128) stub_net_kernel(const float* __restrict__ nn_in, int64_t n, const int64_t* __restrict__ coef,
                                                      int dyadic, float* __restrict__ pi, int64_t pi_stride, float* __restrict__ v) {
  const int64_t g = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (g < n) body_stub(nn_in + g * 196, coef, dyadic, pi + g * pi_stride, v + g);
}
#endif

extern "C" int ifx_stub_policy_value(const float* nn_in, int64_t n, const int64_t* coef, int dyadic, float* pi,
                                     int64_t pi_stride, float* v, void* stream) {
  if (n < 0 || (n > 0 && (!nn_in || !coef || !pi || !v)) || pi_stride < IFX_ACTIONS) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  for (int64_t g = 0; g < n; ++g) body_stub(nn_in + g * 196, coef, dyadic, pi + g * pi_stride, v + g);
  return ifx_after_launch();
#else
  stub_net_kernel<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(nn_in, n, coef, dyadic, pi, pi_stride, v);
  return ifx_after_launch();
#endif
}

// Rules kernels (one thread per state) and their C ABI.  See include/ifx_b200.h for the contract.
#include "ifx_rules.cuh"
#include "ifx_host.cuh"
#include "ifx_philox.cuh"

// ---------------------------------------------------------------------------------------------
// per-state cores
// ---------------------------------------------------------------------------------------------
IFX_HD void rules_masks_of(const ifx_state& s, uint64_t* m7) {
  const uint64_t own = ifx_own(s);
#pragma unroll
  for (int m = 0; m < 6; ++m) m7[m] = own;
  m7[6] = ifx_spawn_mask(s);
}

IFX_HD int rules_step_of(const ifx_state& in, int action, ifx_state& out) {
  out = in;
  if (in.outcome != IFX_ONGOING || !ifx_action_valid(in, action)) return IFX_F_BADACTION;
  ifx_apply(out, action);
  return 0;
}

IFX_HD void rules_info_of(const ifx_state& s, int32_t* info) {
  const uint64_t own = ifx_own(s), opp = ifx_opp(s);
  const int own_pw = ifx_popc64(s.p0 & own) + 2 * ifx_popc64(s.p1 & own) + 4 * ifx_popc64(s.p2 & own);
  const int tot = ifx_total_power(s);
  info[0] = ifx_popc64(own) - ifx_popc64(opp);     // score = piece_count_diff (InflexionGame.py:198-199, 319-323)
  info[1] = 2 * own_pw - tot;                      // power_diff for the side to move (:312-317)
  info[2] = tot;                                   // total_power (:347)
  info[3] = ifx_num_valid(s);
}

// index of the k-th (0-based) set bit of m
IFX_HD int rules_nth_bit(uint64_t m, int k) {
  while (k-- > 0) m &= m - 1;
  int idx = 0;
  while (!((m >> idx) & 1ull)) ++idx;
  return idx;
}

// uniform-random valid action, drawn as RandomPlayer does (InflexionPlayers.py:24-34: choice over the valid
// actions in ascending order): the (w0 % count)-th valid action, Philox counter = (stream, step, 0, 3).
IFX_HD int rules_random_action_of(const ifx_state& s, uint64_t seed, uint32_t stream, uint32_t step) {
  if (s.outcome != IFX_ONGOING) return -1;
  const uint64_t own = ifx_own(s), spawn = ifx_spawn_mask(s);
  const int n_own = ifx_popc64(own), n_sp = ifx_popc64(spawn);
  const int count = 6 * n_own + n_sp;
  if (count == 0) return -1;
  const IfxRand4 r = ifx_philox(seed, stream, step, 0u, 3u);
  int k = (int)(r.w[0] % (uint32_t)count);
  if (k < 6 * n_own) { const int m = k / n_own; return m * IFX_CELLS + rules_nth_bit(own, k - m * n_own); }
  return IFX_SPAWN * IFX_CELLS + rules_nth_bit(spawn, k - 6 * n_own);
}

// GreedyPlayer.play (InflexionPlayers.py:61-77): one-ply look-ahead on score() = #own - #opponent stacks from the
// mover's side; candidates are sorted (score, action) descending, so ties go to the HIGHEST action index.
IFX_HD int rules_greedy_action_of(const ifx_state& s) {
  if (s.outcome != IFX_ONGOING) return -1;
  const uint64_t own = ifx_own(s), spawn = ifx_spawn_mask(s);
  int best = -1, best_score = -1000;
  for (int a = 0; a < IFX_ACTIONS; ++a) {
    const int m = a / IFX_CELLS, cell = a - m * IFX_CELLS;
    if (!(((m == IFX_SPAWN ? spawn : own) >> cell) & 1ull)) continue;
    ifx_state t = s;
    ifx_apply(t, a);
    // after the move the side to move is the opponent: mover's score = #opp(t) - #own(t)
    const int score = ifx_popc64(ifx_opp(t)) - ifx_popc64(ifx_own(t));
    if (score >= best_score) { best_score = score; best = a; }
  }
  return best;
}

// value of plane c at (symmetry-mapped) cell, as float
IFX_HD int rules_plane_value(const ifx_state& s, int c, int src_cell) {
  switch (c) {
    case 0: return (int)((ifx_own(s) >> src_cell) & 1ull);
    case 1: return (int)((ifx_opp(s) >> src_cell) & 1ull);
    case 2: return (int)s.turn;
    default: return ifx_can_spawn(s) ? 1 : 0;
  }
}

#ifndef IFX_HOSTSIM
// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
#define RULES_BLOCK 128

// 16-byte chunk c (0..3) of state t inside the staging tile; XOR swizzle makes both the coalesced
// tile copy and the per-thread 64-byte reads bank-conflict free.
__device__ __forceinline__ int tile_chunk(int t, int c) { return 4 * t + (c ^ ((t >> 1) & 3)); }

__global__ void __launch_bounds__(RULES_BLOCK)
rules_mask_step_kernel(const uint4* __restrict__ in, const int32_t* __restrict__ actions, int64_t n,
                       uint64_t* __restrict__ masks, uint4* __restrict__ out, int32_t* __restrict__ status) {
  __shared__ uint4 tile[RULES_BLOCK * 4];                       // 8 KB: states in, then states out
  __shared__ __align__(16) uint64_t mtile[RULES_BLOCK * 7];     // 7 KB: packed masks
  const int t = threadIdx.x;
  const int64_t base = (int64_t)blockIdx.x * RULES_BLOCK;
  const int cnt = (int)min((int64_t)RULES_BLOCK, n - base);
  // coalesced 128-bit loads of the block's states
  const uint4* src = in + base * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int u = j * RULES_BLOCK + t;
    if (u < cnt * 4) tile[tile_chunk(u >> 2, u & 3)] = __ldcs(src + u);
  }
  __syncthreads();
  union { ifx_state s; uint4 q[4]; } a, b;
  int st = 0;
  if (t < cnt) {
#pragma unroll
    for (int c = 0; c < 4; ++c) a.q[c] = tile[tile_chunk(t, c)];
    if (masks) rules_masks_of(a.s, &mtile[t * 7]);
    if (actions) st = rules_step_of(a.s, __ldcs(actions + base + t), b.s);
  }
  __syncthreads();
  if (actions) {
    if (t < cnt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tile[tile_chunk(t, c)] = b.q[c];
      if (status) status[base + t] = st;
    }
    __syncthreads();
    uint4* dst = out + base * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int u = j * RULES_BLOCK + t;
      if (u < cnt * 4) __stcs(dst + u, tile[tile_chunk(u >> 2, u & 3)]);
    }
  }
  if (masks) {
    // block's masks are 128*56 B contiguous and 16-B aligned (base*56 is a multiple of 16)
    const int n16 = cnt * 7 / 2;                                // whole 16-byte pieces
    uint4* mdst = reinterpret_cast<uint4*>(masks + base * 7);
    const uint4* msrc = reinterpret_cast<const uint4*>(mtile);
    for (int u = t; u < n16; u += RULES_BLOCK) __stcs(mdst + u, msrc[u]);
    if ((cnt * 7) & 1) { if (t == 0) masks[base * 7 + cnt * 7 - 1] = mtile[cnt * 7 - 1]; }
  }
}

template <typename T>
__global__ void rules_planes_kernel(const ifx_state* __restrict__ states, const int32_t* __restrict__ sym,
                                    int64_t n, T* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * 196) return;
  const int64_t i = e / 196;
  const int rem = (int)(e - i * 196), c = rem / IFX_CELLS, cell = rem - c * IFX_CELLS;
  const int sy = sym ? __ldg(sym + i) : 0;
  union { ifx_state s; uint4 q[3]; } a;                         // first 48 bytes hold everything needed
  const uint4* p = reinterpret_cast<const uint4*>(states + i);
  a.q[0] = __ldg(p); a.q[1] = __ldg(p + 1); a.q[2] = __ldg(p + 2);
  out[e] = (T)rules_plane_value(a.s, c, c < 2 ? ifx_sym_src(sy, cell) : cell);
}

template <typename T>
__global__ void rules_symmetries_kernel(const T* __restrict__ in, int64_t n, int channels, T* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)36 * channels * IFX_CELLS;
  if (e >= n * per) return;
  const int64_t i = e / per;
  int rem = (int)(e - i * per);
  const int f = rem / (channels * IFX_CELLS); rem -= f * channels * IFX_CELLS;
  const int c = rem / IFX_CELLS, cell = rem - c * IFX_CELLS;
  out[e] = __ldg(in + (i * channels + c) * IFX_CELLS + ifx_sym_src(ifx_sym36(f), cell));
}

template <typename T>
__global__ void rules_transform_kernel(const T* __restrict__ in, const int32_t* __restrict__ sym, int64_t n,
                                       int channels, T* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)channels * IFX_CELLS;
  if (e >= n * per) return;
  const int64_t i = e / per;
  const int rem = (int)(e - i * per), c = rem / IFX_CELLS, cell = rem - c * IFX_CELLS;
  out[e] = __ldg(in + (i * channels + c) * IFX_CELLS + ifx_sym_src(__ldg(sym + i), cell));
}

__global__ void rules_random_actions_kernel(const ifx_state* __restrict__ states, int64_t n, uint64_t seed,
                                            uint32_t step, int32_t* __restrict__ actions) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  union { ifx_state s; uint4 q[4]; } a;
  const uint4* p = reinterpret_cast<const uint4*>(states + i);
#pragma unroll
  for (int c = 0; c < 4; ++c) a.q[c] = __ldg(p + c);
  actions[i] = rules_random_action_of(a.s, seed, a.s.tag, step);
}

__global__ void rules_greedy_actions_kernel(const ifx_state* __restrict__ states, int64_t n, int32_t* __restrict__ actions) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  union { ifx_state s; uint4 q[4]; } a;
  const uint4* p = reinterpret_cast<const uint4*>(states + i);
#pragma unroll
  for (int c = 0; c < 4; ++c) a.q[c] = __ldg(p + c);
  actions[i] = rules_greedy_action_of(a.s);
}

__global__ void rules_info_kernel(const ifx_state* __restrict__ states, int64_t n, int32_t* __restrict__ info) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  union { ifx_state s; uint4 q[4]; } a;
  const uint4* p = reinterpret_cast<const uint4*>(states + i);
#pragma unroll
  for (int c = 0; c < 4; ++c) a.q[c] = __ldg(p + c);
  int32_t v[4];
  rules_info_of(a.s, v);
  *reinterpret_cast<int4*>(info + i * 4) = make_int4(v[0], v[1], v[2], v[3]);
}
#endif  // !IFX_HOSTSIM

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int ifx_rules_mask_step(const ifx_state* in, const int32_t* actions, int64_t n, uint64_t* masks,
                        ifx_state* out, int32_t* status, void* stream) {
  if (n < 0) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  if (!in || (actions && !out) || (!actions && !masks)) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  for (int64_t i = 0; i < n; ++i) {
    if (masks) rules_masks_of(in[i], masks + i * 7);
    if (actions) { int st = rules_step_of(in[i], actions[i], out[i]); if (status) status[i] = st; }
  }
  return IFX_OK;
#else
  if (((uintptr_t)in | (uintptr_t)out | (uintptr_t)masks) & 15) return IFX_E_ARG;
  const int64_t blocks = (n + RULES_BLOCK - 1) / RULES_BLOCK;
  rules_mask_step_kernel<<<(unsigned)blocks, RULES_BLOCK, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(in), actions, n, masks, reinterpret_cast<uint4*>(out), status);
  return ifx_after_launch();
#endif
}

int ifx_rules_planes(const ifx_state* states, const int32_t* sym, int64_t n, int dtype, void* out, void* stream) {
  if (n < 0 || (n > 0 && (!states || !out)) || (dtype != 0 && dtype != 1)) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  for (int64_t i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c)
      for (int cell = 0; cell < IFX_CELLS; ++cell) {
        const int v = rules_plane_value(states[i], c, c < 2 ? ifx_sym_src(sym ? sym[i] : 0, cell) : cell);
        if (dtype == 0) ((float*)out)[i * 196 + c * 49 + cell] = (float)v;
        else ((int64_t*)out)[i * 196 + c * 49 + cell] = v;
      }
  return IFX_OK;
#else
  const int64_t total = n * 196;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == 0) rules_planes_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(states, sym, n, (float*)out);
  else rules_planes_kernel<long long><<<blocks, 256, 0, (cudaStream_t)stream>>>(states, sym, n, (long long*)out);
  return ifx_after_launch();
#endif
}

int ifx_rules_symmetries(const void* in, int64_t n, int channels, int elem_size, void* out, void* stream) {
  if (n < 0 || channels <= 0 || (elem_size != 4 && elem_size != 8) || (n > 0 && (!in || !out))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  const int64_t per = (int64_t)36 * channels * IFX_CELLS;
#ifdef IFX_HOSTSIM
  for (int64_t e = 0; e < n * per; ++e) {
    const int64_t i = e / per;
    int rem = (int)(e - i * per);
    const int f = rem / (channels * IFX_CELLS); rem -= f * channels * IFX_CELLS;
    const int c = rem / IFX_CELLS, cell = rem - c * IFX_CELLS;
    const int64_t s = (i * channels + c) * IFX_CELLS + ifx_sym_src(ifx_sym36(f), cell);
    if (elem_size == 4) ((uint32_t*)out)[e] = ((const uint32_t*)in)[s];
    else ((uint64_t*)out)[e] = ((const uint64_t*)in)[s];
  }
  return IFX_OK;
#else
  const unsigned blocks = (unsigned)((n * per + 255) / 256);
  if (elem_size == 4)
    rules_symmetries_kernel<uint32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)in, n, channels, (uint32_t*)out);
  else
    rules_symmetries_kernel<unsigned long long><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)in, n, channels, (unsigned long long*)out);
  return ifx_after_launch();
#endif
}

int ifx_rules_transform(const void* in, const int32_t* sym, int64_t n, int channels, int elem_size, void* out,
                        void* stream) {
  if (n < 0 || channels <= 0 || (elem_size != 4 && elem_size != 8) || (n > 0 && (!in || !out || !sym))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  const int64_t per = (int64_t)channels * IFX_CELLS;
#ifdef IFX_HOSTSIM
  for (int64_t e = 0; e < n * per; ++e) {
    const int64_t i = e / per;
    const int rem = (int)(e - i * per), c = rem / IFX_CELLS, cell = rem - c * IFX_CELLS;
    if (sym[i] < 0 || sym[i] >= 126) return IFX_E_ARG;
    const int64_t s = (i * channels + c) * IFX_CELLS + ifx_sym_src(sym[i], cell);
    if (elem_size == 4) ((uint32_t*)out)[e] = ((const uint32_t*)in)[s];
    else ((uint64_t*)out)[e] = ((const uint64_t*)in)[s];
  }
  return IFX_OK;
#else
  const unsigned blocks = (unsigned)((n * per + 255) / 256);
  if (elem_size == 4)
    rules_transform_kernel<uint32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)in, sym, n, channels, (uint32_t*)out);
  else
    rules_transform_kernel<unsigned long long><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)in, sym, n, channels, (unsigned long long*)out);
  return ifx_after_launch();
#endif
}

int ifx_rules_random_actions(const ifx_state* states, int64_t n, uint64_t seed, uint32_t step, int32_t* actions,
                             void* stream) {
  if (n < 0 || (n > 0 && (!states || !actions))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  for (int64_t i = 0; i < n; ++i) actions[i] = rules_random_action_of(states[i], seed, states[i].tag, step);
  return IFX_OK;
#else
  rules_random_actions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(states, n, seed, step, actions);
  return ifx_after_launch();
#endif
}

int ifx_rules_greedy_actions(const ifx_state* states, int64_t n, int32_t* actions, void* stream) {
  if (n < 0 || (n > 0 && (!states || !actions))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  for (int64_t i = 0; i < n; ++i) actions[i] = rules_greedy_action_of(states[i]);
  return IFX_OK;
#else
  rules_greedy_actions_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(states, n, actions);
  return ifx_after_launch();
#endif
}

int ifx_rules_info(const ifx_state* states, int64_t n, int32_t* info, void* stream) {
  if (n < 0 || (n > 0 && (!states || !info))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
#ifdef IFX_HOSTSIM
  for (int64_t i = 0; i < n; ++i) rules_info_of(states[i], info + i * 4);
  return IFX_OK;
#else
  rules_info_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(states, n, info);
  return ifx_after_launch();
#endif
}

// ---- host-buffer forms ----------------------------------------------------------------------
int ifx_rules_mask_step_host(const ifx_state* in, const int32_t* actions, int64_t n, uint64_t* masks,
                             ifx_state* out, int32_t* status) {
  if (n < 0 || (n > 0 && !in)) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  IfxScratch sc;
  ifx_state* d_in = (ifx_state*)sc.take(n * sizeof(ifx_state));
  ifx_state* d_out = actions ? (ifx_state*)sc.take(n * sizeof(ifx_state)) : nullptr;
  uint64_t* d_masks = masks ? (uint64_t*)sc.take(n * 7 * sizeof(uint64_t)) : nullptr;
  int32_t* d_act = actions ? (int32_t*)sc.take(n * sizeof(int32_t)) : nullptr;
  int32_t* d_st = (actions && status) ? (int32_t*)sc.take(n * sizeof(int32_t)) : nullptr;
  if (!sc.ok()) return IFX_E_CUDA;
  if (ifx_h2d(d_in, in, n * sizeof(ifx_state))) return IFX_E_CUDA;
  if (actions && ifx_h2d(d_act, actions, n * sizeof(int32_t))) return IFX_E_CUDA;
  int rc = ifx_rules_mask_step(d_in, d_act, n, d_masks, d_out, d_st, nullptr);
  if (rc) return rc;
  if (actions && ifx_d2h(out, d_out, n * sizeof(ifx_state))) return IFX_E_CUDA;
  if (masks && ifx_d2h(masks, d_masks, n * 7 * sizeof(uint64_t))) return IFX_E_CUDA;
  if (d_st && ifx_d2h(status, d_st, n * sizeof(int32_t))) return IFX_E_CUDA;
  return ifx_sync();
}

int ifx_rules_planes_host(const ifx_state* states, const int32_t* sym, int64_t n, int dtype, void* out) {
  if (n < 0 || (n > 0 && (!states || !out)) || (dtype != 0 && dtype != 1)) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  const size_t ob = (size_t)n * 196 * (dtype ? 8 : 4);
  IfxScratch sc;
  ifx_state* d_in = (ifx_state*)sc.take(n * sizeof(ifx_state));
  int32_t* d_sym = sym ? (int32_t*)sc.take(n * sizeof(int32_t)) : nullptr;
  void* d_out = sc.take(ob);
  if (!sc.ok()) return IFX_E_CUDA;
  if (ifx_h2d(d_in, states, n * sizeof(ifx_state))) return IFX_E_CUDA;
  if (sym && ifx_h2d(d_sym, sym, n * sizeof(int32_t))) return IFX_E_CUDA;
  int rc = ifx_rules_planes(d_in, d_sym, n, dtype, d_out, nullptr);
  if (rc) return rc;
  if (ifx_d2h(out, d_out, ob)) return IFX_E_CUDA;
  return ifx_sync();
}

int ifx_rules_symmetries_host(const void* in, int64_t n, int channels, int elem_size, void* out) {
  if (n < 0 || channels <= 0 || (elem_size != 4 && elem_size != 8) || (n > 0 && (!in || !out))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  const size_t ib = (size_t)n * channels * IFX_CELLS * elem_size;
  IfxScratch sc;
  void* d_in = sc.take(ib);
  void* d_out = sc.take(ib * 36);
  if (!sc.ok()) return IFX_E_CUDA;
  if (ifx_h2d(d_in, in, ib)) return IFX_E_CUDA;
  int rc = ifx_rules_symmetries(d_in, n, channels, elem_size, d_out, nullptr);
  if (rc) return rc;
  if (ifx_d2h(out, d_out, ib * 36)) return IFX_E_CUDA;
  return ifx_sync();
}

int ifx_rules_transform_host(const void* in, const int32_t* sym, int64_t n, int channels, int elem_size, void* out) {
  if (n < 0 || channels <= 0 || (elem_size != 4 && elem_size != 8) || (n > 0 && (!in || !out || !sym))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  for (int64_t i = 0; i < n; ++i) if (sym[i] < 0 || sym[i] >= 126) return IFX_E_ARG;
  const size_t ib = (size_t)n * channels * IFX_CELLS * elem_size;
  IfxScratch sc;
  void* d_in = sc.take(ib);
  void* d_out = sc.take(ib);
  int32_t* d_sym = (int32_t*)sc.take(n * sizeof(int32_t));
  if (!sc.ok()) return IFX_E_CUDA;
  if (ifx_h2d(d_in, in, ib) || ifx_h2d(d_sym, sym, n * sizeof(int32_t))) return IFX_E_CUDA;
  int rc = ifx_rules_transform(d_in, d_sym, n, channels, elem_size, d_out, nullptr);
  if (rc) return rc;
  if (ifx_d2h(out, d_out, ib)) return IFX_E_CUDA;
  return ifx_sync();
}

int ifx_rules_info_host(const ifx_state* states, int64_t n, int32_t* info) {
  if (n < 0 || (n > 0 && (!states || !info))) return IFX_E_ARG;
  if (n == 0) return IFX_OK;
  IfxScratch sc;
  ifx_state* d_in = (ifx_state*)sc.take(n * sizeof(ifx_state));
  int32_t* d_info = (int32_t*)sc.take(n * 16);
  if (!sc.ok()) return IFX_E_CUDA;
  if (ifx_h2d(d_in, states, n * sizeof(ifx_state))) return IFX_E_CUDA;
  int rc = ifx_rules_info(d_in, n, d_info, nullptr);
  if (rc) return rc;
  if (ifx_d2h(info, d_info, n * 16)) return IFX_E_CUDA;
  return ifx_sync();
}

}  // extern "C"

// Tree kernels (one warp per game) and the ifx_mcts_* C ABI.  See include/ifx_b200.h for the contract.
#include "ifx_tree.cuh"
#include "ifx_tree_vl.cuh"
#include "ifx_host.cuh"
#include <new>

#define IFX_TIMING_RING 4096
struct ifx_mcts {
  ifx_mcts_config cfg;
  ifx_mcts_layout lay;
  TreeParams P;
  void* workspace;
  // optional CUDA-event timing of the wave kernel (bench evidence): a ring of event pairs recorded right around
  // the launch on the launching stream, summed by ifx_mcts_wave_time
  bool timing = false;
  int n_timed = 0;
#ifndef IFX_HOSTSIM
  cudaEvent_t ev[IFX_TIMING_RING][2];
  int ev_made = 0;
#endif
};

static int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }
static int32_t sym_cap_of(int32_t num_sims) { const int32_t n = num_sims > 128 ? num_sims : 128; return (n + 31) / 32 * 32; }
static int32_t leaves_of(const ifx_mcts_config* c) { return c->leaves_per_wave > 1 ? c->leaves_per_wave : 1; }
static int32_t path_stride_of(int32_t max_depth) { return (max_depth + IFX_PATH_CHUNK - 1) / IFX_PATH_CHUNK * IFX_PATH_CHUNK; }

static int check_cfg(const ifx_mcts_config* c) {
  if (!c) return IFX_E_ARG;
  if (c->n_games <= 0 || c->nodes_per_game < 32 || c->nodes_per_game > 16383) return IFX_E_ARG;
  if (c->hash_per_game < 64 || (c->hash_per_game & (c->hash_per_game - 1))) return IFX_E_ARG;
  if (c->hash_per_game < 2 * c->nodes_per_game) return IFX_E_ARG;
  if (c->max_depth < 1 || c->max_depth > IFX_MAX_DEPTH || c->num_sims < 0 || c->num_sims >= IFX_NS_LIMIT) return IFX_E_ARG;
  if (c->policy_is_logits != 0 && c->policy_is_logits != 2) return IFX_E_ARG;
  if (c->terminal_sims_per_wave < 0) return IFX_E_ARG;
  if (c->dirichlet_eps < 0.0f || c->dirichlet_eps > 1.0f || (c->dirichlet_eps > 0.0f && !(c->dirichlet_alpha > 0.0f))) return IFX_E_ARG;
  if (c->leaves_per_wave < -1 || c->leaves_per_wave > IFX_VL_MAX_LEAVES) return IFX_E_ARG;
  if (c->leaves_per_wave > 1 && c->dirichlet_eps > 0.0f) return IFX_E_ARG;      // (root noise: sequential mode only)
  return IFX_OK;
}

// ---------------------------------------------------------------------------------------------
// per-game bodies shared by the device kernels and the host simulation
// ---------------------------------------------------------------------------------------------
IFX_D void body_reset(const TreeParams& P, int g, const ifx_state* roots, const int32_t* which) {
  if (which && !which[g]) return;
  if (IFX_LANE0) P.roots[g] = roots[g];
  IFX_SYNCWARP();
  tree_begin_move(P, g, true);
}

IFX_D void body_set_roots(const TreeParams& P, int g, const ifx_state* roots) {
  const bool searched = P.roots[g].outcome == IFX_ONGOING;   // was the previous root one the search worked on?
  IFX_SYNCWARP();
  if (IFX_LANE0) P.roots[g] = roots[g];
  IFX_SYNCWARP();
  tree_begin_move(P, g, false, searched);
}

IFX_D void body_advance(const TreeParams& P, int g, const int32_t* actions) {
  ifx_state root = P.roots[g];
  const int a = actions[g];
  if (root.outcome == IFX_ONGOING) {
    if (!ifx_action_valid(root, a)) {
      if (IFX_LANE0) P.games[g].faults |= IFX_F_BADACTION;
      IFX_SYNCWARP();
      return;
    }
    ifx_apply(root, a);                                      // Coach.py:82
    IFX_SYNCWARP();
    if (IFX_LANE0) P.roots[g] = root;
    IFX_SYNCWARP();
  }
  tree_begin_move(P, g, false);
}

// Coach.learn starts a new episode when one ends (Coach.py:108-112): finished games get the start position, an
// empty tree and a fresh RNG stream (tag + tag_stride); `count` (nullable) accumulates how many were restarted.
IFX_D void body_restart_finished(const TreeParams& P, int g, uint32_t tag_stride, int32_t* count) {
  const ifx_state old = P.roots[g];
  if (old.outcome == IFX_ONGOING) return;
  IFX_SYNCWARP();
  if (IFX_LANE0) {
    ifx_state s;
    memset(&s, 0, sizeof(s));
    s.max_turns = old.max_turns;
    s.tag = old.tag + tag_stride;
    P.roots[g] = s;
#ifdef IFX_HOSTSIM
    if (count) *count += 1;
#else
    if (count) atomicAdd(count, 1);
#endif
  }
  IFX_SYNCWARP();
  tree_begin_move(P, g, true);
}

IFX_D void body_root_stats(const TreeParams& P, int g, int32_t* counts, float* q32, double* q64, uint8_t* qkind,
                           int32_t* ns) {
  const int slot = tree_root_slot(P, g);
  const uint32_t* nrow = slot >= 0 ? reinterpret_cast<const uint32_t*>(tree_row(P, g, slot, 1)) : nullptr;
  const float* qrow = slot >= 0 ? tree_row(P, g, slot, 2) : nullptr;
  const double* qd = slot >= 0 ? P.q64 + ((int64_t)g * P.C + slot) * 344 : nullptr;
  const ifx_state root = P.roots[g];
  const ValidSet V = ifx_valid_set(ifx_own(root), ifx_spawn_mask(root));   // rows are compact: entry = rank among valid actions
  IFX_FOR_LANES(l) {
    for (int a = l; a < IFX_ACTIONS; a += 32) {
      const bool valid = nrow && tree_action_bit(V.own, V.spawn, a);
      const int ci = valid ? ifx_valid_below(V, a) : 0;
      const uint32_t raw = valid ? nrow[ci] : 0u;
      const uint32_t n = raw & IFX_N_COUNT;
      const bool is64 = (raw & IFX_N_TAG) != 0;
      const int64_t o = (int64_t)g * IFX_ACTIONS + a;
      if (counts) counts[o] = (int32_t)n;
      if (q32) q32[o] = n ? qrow[ci] : 0.0f;
      if (q64) q64[o] = n ? (is64 ? qd[ci] : (double)qrow[ci]) : 0.0;
      if (qkind) qkind[o] = (uint8_t)(n && !is64);
    }
  }
  if (ns && IFX_LANE0) ns[g] = slot >= 0 ? P.node_head[(int64_t)g * P.C + slot].ns : 0;
}

// optional Dirichlet root noise for roots expanded by the last wave (see tree_root_noise; ifx_mcts_config)
IFX_D void body_root_noise(const TreeParams& P, int g) {
  IfxGame G = P.games[g];
  if (!G.noise_due) return;
  tree_root_noise(P, g, G, G.root_slot, G.root_nv);
  if (IFX_LANE0) P.games[g].noise_due = 0;
}

IFX_D void body_status(const TreeParams& P, int g, int32_t* status) {
  if (!IFX_LANE0) return;
  const IfxGame G = P.games[g];
  int32_t* s = status + (int64_t)g * 8;
  s[0] = G.sims_done; s[1] = G.pending; s[2] = G.faults; s[3] = G.move_idx;
  s[4] = P.roots[g].outcome; s[5] = G.n_live; s[6] = G.free_top; s[7] = G.path_len;
}

#ifndef IFX_HOSTSIM
// One warp per CTA: a game's slot is handed to the next game the moment its simulation ends (with four warps per
// CTA the three siblings of a deep game idle until it finishes).  Compiled for 32 resident CTAs per SM (the hardware
// limit; 56 registers, no spills; ~6 KB of shared memory each): 4 736 warp slots, so the 4 096 games of the bench
// workload run as ONE generation.
#ifndef TREE_WARPS
#define TREE_WARPS 1
#endif
#define TREE_BLOCK (TREE_WARPS * 32)
#ifndef IFX_WAVE_MIN_BLOCKS
#define IFX_WAVE_MIN_BLOCKS 32     // resident CTAs per SM the wave kernel is compiled for (register cap 65536/(32*n))
#endif

__global__ void tree_sym_lut_kernel() {                      // <<<126, 64>>>: one block per symmetry
  g_ifx_sym_lut[blockIdx.x * 64 + threadIdx.x] = threadIdx.x < IFX_CELLS ? (uint8_t)ifx_sym_src(blockIdx.x, threadIdx.x) : 0;
}

template <bool DBG>
__global__ void __launch_bounds__(TREE_BLOCK, IFX_WAVE_MIN_BLOCKS)
tree_wave_kernel(TreeParams P, const float* __restrict__ nn_pi, int64_t pi_stride, const float* __restrict__ nn_v,
                 float* __restrict__ nn_in, int g_first, int g_end) {
  __shared__ ExpandScratch scratch[TREE_WARPS];
  __shared__ RootMirror mirror[TREE_WARPS];
  const int w = threadIdx.x >> 5, g = g_first + blockIdx.x * TREE_WARPS + w;
  if (g < g_end) tree_wave_game<DBG>(P, g, nn_pi, pi_stride, nn_v, nn_in, &scratch[w], &mirror[w]);
}

// leaf-parallel mode (performance mode, ifx_tree_vl.cuh): one warp per game as well, no register ceiling to meet
__global__ void __launch_bounds__(32)
tree_wave_vl_kernel(TreeParams P, const float* __restrict__ nn_pi, int64_t pi_stride, const float* __restrict__ nn_v,
                    float* __restrict__ nn_in, int g_first, int g_end) {
  __shared__ ExpandScratch scratch;
  __shared__ VlScratch vl;
  const int g = g_first + blockIdx.x;
  if (g < g_end) tree_wave_game_vl(P, g, nn_pi, pi_stride, nn_v, nn_in, &scratch, &vl);
}

__global__ void __launch_bounds__(TREE_BLOCK)
tree_noise_kernel(TreeParams P, int g_first, int g_end) {
  const int g = g_first + blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
  if (g < g_end) body_root_noise(P, g);
}
__global__ void __launch_bounds__(TREE_BLOCK)
tree_reset_kernel(TreeParams P, const ifx_state* __restrict__ roots, const int32_t* __restrict__ which) {
  const int g = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
  if (g < P.G) body_reset(P, g, roots, which);
}
__global__ void __launch_bounds__(TREE_BLOCK)
tree_set_roots_kernel(TreeParams P, const ifx_state* __restrict__ roots) {
  const int g = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
  if (g < P.G) body_set_roots(P, g, roots);
}
__global__ void __launch_bounds__(TREE_BLOCK)
tree_advance_kernel(TreeParams P, const int32_t* __restrict__ actions) {
  const int g = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
  if (g < P.G) body_advance(P, g, actions);
}
__global__ void __launch_bounds__(TREE_BLOCK)
tree_restart_kernel(TreeParams P, uint32_t tag_stride, int32_t* count) {
  const int g = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
  if (g < P.G) body_restart_finished(P, g, tag_stride, count);
}
__global__ void __launch_bounds__(TREE_BLOCK)
tree_root_stats_kernel(TreeParams P, int32_t* counts, float* q32, double* q64, uint8_t* qkind, int32_t* ns) {
  const int g = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
  if (g < P.G) body_root_stats(P, g, counts, q32, q64, qkind, ns);
}
__global__ void __launch_bounds__(TREE_BLOCK)
tree_sample_kernel(TreeParams P, int temp_threshold, int32_t* __restrict__ actions) {
  __shared__ int buf[TREE_WARPS][IFX_ROW];
  __shared__ double dbuf[TREE_WARPS][IFX_ROW];
  const int w = threadIdx.x >> 5, g = blockIdx.x * TREE_WARPS + w;
  if (g >= P.G) return;
  const int a = tree_sample_action(P, g, temp_threshold, buf[w], dbuf[w]);
  if (IFX_LANE0) actions[g] = a;
}
__global__ void __launch_bounds__(TREE_BLOCK)
tree_status_kernel(TreeParams P, int32_t* __restrict__ status) {
  const int g = blockIdx.x * TREE_WARPS + (threadIdx.x >> 5);
  if (g < P.G) body_status(P, g, status);
}
__global__ void tree_get_roots_kernel(TreeParams P, uint4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)P.G * 4) out[i] = reinterpret_cast<const uint4*>(P.roots)[i];
}
__global__ void tree_counters_kernel(TreeParams P, unsigned long long* __restrict__ out) {
  __shared__ unsigned long long acc[4];
  if (threadIdx.x < 4) acc[threadIdx.x] = 0ull;
  __syncthreads();
  unsigned long long v = 0, e = 0, t = 0, s = 0;
  for (int g = threadIdx.x; g < P.G; g += blockDim.x) {
    const IfxGame G = P.games[g];
    v += G.n_visits; e += G.n_expand; t += G.n_term; s += G.n_sims;
  }
  atomicAdd(&acc[0], v); atomicAdd(&acc[1], e); atomicAdd(&acc[2], t); atomicAdd(&acc[3], s);
  __syncthreads();
  if (threadIdx.x < 4) out[threadIdx.x] = acc[threadIdx.x];
}
// games that will search in the coming move (root ongoing, simulations left, no fault), ascending; one block, two passes
__global__ void __launch_bounds__(1024) tree_live_rows_kernel(TreeParams P, int32_t* __restrict__ live_idx, int32_t* __restrict__ count) {
  __shared__ int32_t part[1024];
  const int t = threadIdx.x, per = (P.G + 1023) / 1024, lo = t * per, hi = min(P.G, lo + per);
  int n = 0;
  for (int g = lo; g < hi; ++g) {
    const IfxGame G = P.games[g];
    n += (P.roots[g].outcome == IFX_ONGOING && G.sims_done < P.num_sims && !G.faults) ? 1 : 0;
  }
  part[t] = n;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {                 // inclusive scan (Hillis-Steele)
    const int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int pos = part[t] - n;
  for (int g = lo; g < hi; ++g) {
    const IfxGame G = P.games[g];
    if (P.roots[g].outcome == IFX_ONGOING && G.sims_done < P.num_sims && !G.faults) live_idx[pos++] = g;
  }
  if (t == 1023) *count = part[1023];
}
#define TREE_GRID(P) ((unsigned)(((P).G + TREE_WARPS - 1) / TREE_WARPS))
#endif

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int ifx_mcts_get_layout(const ifx_mcts_config* c, ifx_mcts_layout* L) {
  if (check_cfg(c) || !L) return IFX_E_ARG;
  const int64_t G = c->n_games, C = c->nodes_per_game, H = c->hash_per_game;
  int64_t off = 0;
  L->rows = off;       off = align256(off + G * C * 3 * IFX_ROW * 4);
  L->q64 = off;        off = align256(off + G * C * 344 * 8);
  L->node_key = off;   off = align256(off + G * C * 16);
  L->node_head = off;  off = align256(off + G * C * 16);
  L->hash = off;       off = align256(off + G * H * 16);
  L->free_stack = off; off = align256(off + G * C * 2);
  L->per_game = off;   off = align256(off + G * (int64_t)sizeof(IfxGame));
  L->path = off;       off = align256(off + G * (int64_t)leaves_of(c) * path_stride_of(c->max_depth) * 16);
  L->roots = off;      off = align256(off + G * (int64_t)sizeof(ifx_state));
  L->leaves = off;     off = align256(off + G * (int64_t)leaves_of(c) * (int64_t)sizeof(ifx_state));
  L->sym = off;        off = align256(off + G * (int64_t)sym_cap_of(c->num_sims) * 64);
  L->total = off;
  return IFX_OK;
}

size_t ifx_mcts_workspace_bytes(const ifx_mcts_config* c) {
  ifx_mcts_layout L;
  return ifx_mcts_get_layout(c, &L) ? 0 : (size_t)L.total;
}

int ifx_mcts_create(const ifx_mcts_config* c, void* ws, size_t bytes, ifx_mcts** out) {
  ifx_mcts_layout L;
  if (!out || !ws || ifx_mcts_get_layout(c, &L)) return IFX_E_ARG;
  if (bytes < (size_t)L.total || ((uintptr_t)ws & 255)) return IFX_E_ARG;
  ifx_mcts* m = new (std::nothrow) ifx_mcts;
  if (!m) return IFX_E_ARG;
  m->cfg = *c; m->lay = L; m->workspace = ws;
  char* b = (char*)ws;
  TreeParams& P = m->P;
  P.rows = (float*)(b + L.rows); P.q64 = (double*)(b + L.q64); P.node_key = (IfxKey*)(b + L.node_key);
  P.node_head = (IfxHead*)(b + L.node_head); P.hash = (IfxKey*)(b + L.hash);
  P.free_stack = (uint16_t*)(b + L.free_stack); P.games = (IfxGame*)(b + L.per_game);
  P.path = (IfxU4*)(b + L.path); P.roots = (ifx_state*)(b + L.roots); P.leaves = (ifx_state*)(b + L.leaves);
  P.sym = (uint8_t*)(b + L.sym); P.sym_cap = sym_cap_of(c->num_sims);
  P.G = c->n_games; P.C = c->nodes_per_game; P.H = c->hash_per_game; P.maxd = c->max_depth;
  P.path_stride = path_stride_of(c->max_depth); P.leaves_per_wave = leaves_of(c); P.vl_mode = (c->leaves_per_wave > 1 || c->leaves_per_wave == -1) ? 1 : 0;
  P.num_sims = c->num_sims; P.logits = c->policy_is_logits; P.term_budget = c->terminal_sims_per_wave;
  P.dir_alpha = c->dirichlet_alpha; P.dir_eps = c->dirichlet_eps; P.cpuct = c->cpuct; P.seed = c->seed;
  P.debug = nullptr;
#ifndef IFX_HOSTSIM
  // one warp per CTA, ~6 KB of shared memory each: ask for the largest shared-memory carve-out so that the
  // register-limited number of CTAs per SM (IFX_WAVE_MIN_BLOCKS) is also what shared memory allows
  cudaFuncSetAttribute(tree_wave_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(tree_wave_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  {                                                         // spread-line table -> constant memory (idempotent)
    static const int DR[6] = {1, -1, 0, 0, 1, -1}, DQ[6] = {0, 0, 1, -1, -1, 1};   // (dr, dq) of move 0..5, InflexionGame.py:15-21
    static uint64_t lut[6 * IFX_CELLS * 6];
    for (int mv = 0; mv < 6; ++mv)
      for (int cell = 0; cell < IFX_CELLS; ++cell) {
        int r = cell / IFX_N, q = cell % IFX_N;
        uint64_t line = 0;
        for (int k = 1; k <= 6; ++k) {
          r = (r + DR[mv] + IFX_N) % IFX_N; q = (q + DQ[mv] + IFX_N) % IFX_N;
          line |= 1ull << (r * IFX_N + q);
          lut[(mv * IFX_CELLS + cell) * 6 + (k - 1)] = line;
        }
      }
    if (cudaMemcpyToSymbol(g_ifx_line_lut, lut, sizeof(lut)) != cudaSuccess) { delete m; return ifx_set_cuda_error((int)cudaGetLastError()); }
  }
  tree_sym_lut_kernel<<<126, 64>>>();                       // idempotent; ordered before any later launch of this process
  if (cudaDeviceSynchronize() != cudaSuccess) { delete m; return ifx_set_cuda_error((int)cudaGetLastError()); }
#endif
  *out = m;
  return IFX_OK;
}

void ifx_mcts_destroy(ifx_mcts* m) {
  if (!m) return;
#ifndef IFX_HOSTSIM
  for (int i = 0; i < m->ev_made; ++i) { cudaEventDestroy(m->ev[i][0]); cudaEventDestroy(m->ev[i][1]); }
#endif
  delete m;
}

int ifx_mcts_wave_timing(ifx_mcts* m, int enable) {
  if (!m) return IFX_E_ARG;
  m->timing = enable != 0;
  m->n_timed = 0;
  return IFX_OK;
}

int ifx_mcts_wave_time(ifx_mcts* m, double* total_ms, int32_t* launches) {
  if (!m || !total_ms || !launches) return IFX_E_ARG;
  double sum = 0.0;
#ifndef IFX_HOSTSIM
  for (int i = 0; i < m->n_timed; ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(m->ev[i][1]) || cudaEventElapsedTime(&ms, m->ev[i][0], m->ev[i][1]))
      return ifx_set_cuda_error((int)cudaGetLastError());
    sum += ms;
  }
#endif
  *total_ms = sum; *launches = m->n_timed;
  m->n_timed = 0;
  return IFX_OK;
}

int ifx_mcts_set_params(ifx_mcts* m, int32_t num_sims, float cpuct) {
  if (!m || num_sims < 0 || num_sims >= IFX_NS_LIMIT) return IFX_E_ARG;
  m->P.num_sims = m->cfg.num_sims = num_sims;
  m->P.cpuct = m->cfg.cpuct = cpuct;
  return IFX_OK;
}

int ifx_mcts_set_debug(ifx_mcts* m, uint64_t* buf) {
  if (!m) return IFX_E_ARG;
  m->P.debug = buf;
  return IFX_OK;
}

int ifx_mcts_reset_games(ifx_mcts* m, const ifx_state* roots, const int32_t* which, void* stream) {
  if (!m || !roots) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  for (int g = 0; g < m->P.G; ++g) body_reset(m->P, g, roots, which);
  return ifx_after_launch();
#else
  tree_reset_kernel<<<TREE_GRID(m->P), TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, roots, which);
  return ifx_after_launch();
#endif
}

int ifx_mcts_set_roots(ifx_mcts* m, const ifx_state* roots, void* stream) {
  if (!m || !roots) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  for (int g = 0; g < m->P.G; ++g) body_set_roots(m->P, g, roots);
  return ifx_after_launch();
#else
  tree_set_roots_kernel<<<TREE_GRID(m->P), TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, roots);
  return ifx_after_launch();
#endif
}

int ifx_mcts_wave(ifx_mcts* m, const float* nn_pi, int64_t pi_stride, const float* nn_v, float* nn_in, void* stream) {
  return m ? ifx_mcts_wave_part(m, 0, m->P.G, nn_pi, pi_stride, nn_v, nn_in, stream) : IFX_E_ARG;
}

int ifx_mcts_wave_part(ifx_mcts* m, int32_t first_game, int32_t n_games, const float* nn_pi, int64_t pi_stride,
                       const float* nn_v, float* nn_in, void* stream) {
  if (!m || !nn_pi || !nn_in || pi_stride < IFX_ACTIONS) return IFX_E_ARG;
  if (first_game < 0 || n_games < 0 || first_game + n_games > m->P.G) return IFX_E_ARG;
  if (n_games == 0) return IFX_OK;
  const int g_end = first_game + n_games;
  if (m->P.logits == 2) {                  // nn_pi = binary16 head rows: 343 logits + value, 8-byte aligned rows
    if (pi_stride < 344 || (pi_stride & 3) || ((uintptr_t)nn_pi & 7)) return IFX_E_ARG;
  } else if (!nn_v) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  static ExpandScratch scratch;
  static RootMirror mirror;
  static VlScratch vl;
  if (m->P.vl_mode) {
    for (int g = first_game; g < g_end; ++g) tree_wave_game_vl(m->P, g, nn_pi, pi_stride, nn_v, nn_in, &scratch, &vl);
    return ifx_after_launch();
  }
  for (int g = first_game; g < g_end; ++g) tree_wave_game<false>(m->P, g, nn_pi, pi_stride, nn_v, nn_in, &scratch, &mirror);
  if (m->P.dir_eps > 0.0f) for (int g = first_game; g < g_end; ++g) body_root_noise(m->P, g);
  return ifx_after_launch();
#else
  const unsigned grid = (unsigned)((n_games + TREE_WARPS - 1) / TREE_WARPS);
  const bool timed = m->timing && m->n_timed < IFX_TIMING_RING;
  if (timed) {
    while (m->ev_made <= m->n_timed) {
      if (cudaEventCreate(&m->ev[m->ev_made][0]) || cudaEventCreate(&m->ev[m->ev_made][1])) return ifx_set_cuda_error((int)cudaGetLastError());
      m->ev_made += 1;
    }
    cudaEventRecord(m->ev[m->n_timed][0], (cudaStream_t)stream);
  }
  if (m->P.vl_mode) {                        // leaf-parallel performance mode: its own kernel
    tree_wave_vl_kernel<<<(unsigned)n_games, 32, 0, (cudaStream_t)stream>>>(m->P, nn_pi, pi_stride, nn_v, nn_in, first_game, g_end);
    if (timed) { cudaEventRecord(m->ev[m->n_timed][1], (cudaStream_t)stream); m->n_timed += 1; }
    return ifx_after_launch();
  }
  if (m->P.debug)    // per-warp %globaltimer stamps (tools/trace_wave.py): a separate instantiation, so the product kernel carries none of it
    tree_wave_kernel<true><<<grid, TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, nn_pi, pi_stride, nn_v, nn_in, first_game, g_end);
  else
    tree_wave_kernel<false><<<grid, TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, nn_pi, pi_stride, nn_v, nn_in, first_game, g_end);
  if (timed) { cudaEventRecord(m->ev[m->n_timed][1], (cudaStream_t)stream); m->n_timed += 1; }
  if (m->P.dir_eps > 0.0f) {               // noise mode only: one more (tiny) launch per wave
    tree_noise_kernel<<<grid, TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, first_game, g_end);
    ifx_count_launches(1);
  }
  return ifx_after_launch();
#endif
}

int ifx_mcts_root_stats(ifx_mcts* m, int32_t* counts, float* q32, double* q64, uint8_t* qkind, int32_t* ns,
                        void* stream) {
  if (!m) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  for (int g = 0; g < m->P.G; ++g) body_root_stats(m->P, g, counts, q32, q64, qkind, ns);
  return ifx_after_launch();
#else
  tree_root_stats_kernel<<<TREE_GRID(m->P), TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, counts, q32, q64, qkind, ns);
  return ifx_after_launch();
#endif
}

int ifx_mcts_sample_actions(ifx_mcts* m, int32_t temp_threshold, int32_t* actions, void* stream) {
  if (!m || !actions) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  int buf[IFX_ROW];
  double dbuf[IFX_ROW];
  for (int g = 0; g < m->P.G; ++g) actions[g] = tree_sample_action(m->P, g, temp_threshold, buf, dbuf);
  return ifx_after_launch();
#else
  tree_sample_kernel<<<TREE_GRID(m->P), TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, temp_threshold, actions);
  return ifx_after_launch();
#endif
}

int ifx_mcts_advance(ifx_mcts* m, const int32_t* actions, void* stream) {
  if (!m || !actions) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  for (int g = 0; g < m->P.G; ++g) body_advance(m->P, g, actions);
  return ifx_after_launch();
#else
  tree_advance_kernel<<<TREE_GRID(m->P), TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, actions);
  return ifx_after_launch();
#endif
}

int ifx_mcts_restart_finished(ifx_mcts* m, uint32_t tag_stride, int32_t* count, void* stream) {
  if (!m) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  for (int g = 0; g < m->P.G; ++g) body_restart_finished(m->P, g, tag_stride, count);
  return ifx_after_launch();
#else
  tree_restart_kernel<<<TREE_GRID(m->P), TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, tag_stride, count);
  return ifx_after_launch();
#endif
}

int ifx_mcts_status(ifx_mcts* m, int32_t* status, void* stream) {
  if (!m || !status) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  for (int g = 0; g < m->P.G; ++g) body_status(m->P, g, status);
  return ifx_after_launch();
#else
  tree_status_kernel<<<TREE_GRID(m->P), TREE_BLOCK, 0, (cudaStream_t)stream>>>(m->P, status);
  return ifx_after_launch();
#endif
}

int ifx_mcts_get_roots(ifx_mcts* m, ifx_state* roots, void* stream) {
  if (!m || !roots) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  memcpy(roots, m->P.roots, (size_t)m->P.G * sizeof(ifx_state));
  return ifx_after_launch();
#else
  const int64_t n = (int64_t)m->P.G * 4;
  tree_get_roots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(m->P, reinterpret_cast<uint4*>(roots));
  return ifx_after_launch();
#endif
}

int ifx_mcts_live_rows(ifx_mcts* m, int32_t* live_idx, int32_t* count, void* stream) {
  if (!m || !live_idx || !count) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  int32_t n = 0;
  for (int g = 0; g < m->P.G; ++g) {
    const IfxGame& G = m->P.games[g];
    if (m->P.roots[g].outcome == IFX_ONGOING && G.sims_done < m->P.num_sims && !G.faults) live_idx[n++] = g;
  }
  *count = n;
  return ifx_after_launch();
#else
  tree_live_rows_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(m->P, live_idx, count);
  return ifx_after_launch();
#endif
}

int ifx_mcts_counters(ifx_mcts* m, uint64_t* counters, void* stream) {
  if (!m || !counters) return IFX_E_ARG;
#ifdef IFX_HOSTSIM
  counters[0] = counters[1] = counters[2] = counters[3] = 0;
  for (int g = 0; g < m->P.G; ++g) {
    const IfxGame& G = m->P.games[g];
    counters[0] += G.n_visits; counters[1] += G.n_expand; counters[2] += G.n_term; counters[3] += G.n_sims;
  }
  return ifx_after_launch();
#else
  tree_counters_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(m->P, reinterpret_cast<unsigned long long*>(counters));
  return ifx_after_launch();
#endif
}

}  // extern "C"

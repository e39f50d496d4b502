// Tree kernel core: one warp per game tree.  Restates MCTS.search / getActionProb (MCTS.py:33-145) of the
// reference over structure-of-arrays node pools in HBM.  Everything "uniform" (descent, rules step, keys,
// edge arithmetic) is executed redundantly by the 32 lanes; everything over the 343-action row (PUCT
// arg-max, masked renormalisation, root statistics) is split across lanes and combined with shuffles.
//
// Numerics follow SURVEY.md appendix B exactly (numpy 2.x / NEP-50): float32 PUCT with separate
// round-to-nearest operations, numpy's pairwise float32 sum for the prior renormalisation, and edge
// values that are either float32 (came from the net) or Python floats (came from a terminal state).
//
// Latency structure (the kernel is bound by dependent HBM round trips, not by arithmetic): every phase
// issues ALL of its loads before consuming any -- whole P/N/Q rows as 128-bit loads, the pending leaf, its
// policy row, the path's edges -- and state that would cost a dependent load is carried in the per-game
// record instead (root slot, next free slot, hash position of the pending leaf, the path itself, which
// lives one level per lane in registers and is spilled to HBM only while a leaf waits for the net).
#pragma once
#define IFX_WITH_LINE_LUT 1          // (this translation unit owns the constant-memory table of spread lines)
#include "ifx_rules.cuh"
#include "ifx_philox.cuh"

#define IFX_PATH_CHUNK 32          // path levels held in registers at a time: one level per lane
#define IFX_MAX_DEPTH 352          // longest search path accepted (a game has at most 344 plies: no real limit)
#define IFX_NS_LIMIT (1 << 17)     // visit counts occupy 17 bits of an N word; the float32 sqrt identity holds below it

struct alignas(16) IfxGame {       // 64 bytes per game
  int32_t sims_done;               // root-level search() calls finished in the current move (MCTS.py:45-46)
  int32_t pending;                 // 1: `leaf` waits for the net
  int32_t faults;                  // IFX_F_* bits
  int32_t move_idx;                // getActionProb calls so far (RNG coordinate)
  int16_t path_len;                // edges on the path to the pending leaf
  int16_t noise_due;               // 1: the root was expanded in the last wave and still waits for its Dirichlet noise
  int32_t free_top;                // free slots left on the stack (not counting next_slot)
  uint32_t game_id;                // RNG stream coordinate
  int32_t n_live;
  uint32_t n_visits, n_expand, n_term, n_sims;   // counters: V, E, terminal leaves, simulations
  int32_t root_slot;               // node slot of the root, -1 while it is not expanded
  int32_t next_slot;               // slot the next expansion will use (popped ahead of time), -1 = pool empty
  int32_t leaf_hpos;               // hash position (first empty probe slot) found when the pending leaf missed
  int32_t root_nv;                 // number of valid actions of the root (how much of its rows the wave mirrors)
};

struct alignas(16) IfxKey { uint64_t k0, k1; };
struct alignas(16) IfxF4 { float x, y, z, w; };
struct alignas(16) IfxU4 { uint32_t x, y, z, w; };

// Rows are COMPACT: a node stores P, N and Q only for its valid actions, in ascending action order -- entry i of a
// row belongs to the i-th valid action.  The valid set of a node is a pure function of its key (own mask x 6
// spread directions, then the spawn mask: InflexionGame.py:93-100), so nothing but the rows is stored; the
// reference's scan over all 343 actions with `if valids[a]` (MCTS.py:119-120) visits exactly these entries in
// exactly this order.  Under random play a node has 86 valid actions on average, so a visit
// reads 1.5 KB (one 128-entry group of each row) instead of 4.1 KB, and evaluates 4 PUCT terms per lane, not 12.
//
// One 32-bit word per edge in the N row:
//   bits 0..16   visit count N(s,a)              (MCTS.py:27 Nsa)
//   bits 17..30  child hint: node slot the edge led to last time (0x3FFF = none).  Only a hint -- the key of
//                that slot is compared with the recomputed child key before it is trusted, because the
//                reference's table is keyed by planes, not by (s,a)
//   bit 31       Q of this edge is a Python float (kept in q64)
#define IFX_N_TAG 0x80000000u
#define IFX_N_COUNT 0x0001FFFFu
#define IFX_N_HINT_SHIFT 17
#define IFX_N_HINT_NONE 0x3FFFu
#define IFX_N_FRESH (IFX_N_HINT_NONE << IFX_N_HINT_SHIFT)      // unvisited, no hint
#define IFX_GROUP 128                                          // row entries one warp-wide 128-bit load covers

// the valid actions of a node: own stacks (six spread directions each), then the spawn cells
struct ValidSet { uint64_t own, spawn; int n_own, nv; };
IFX_HD ValidSet ifx_valid_set(uint64_t own, uint64_t spawn) {
  ValidSet V; V.own = own; V.spawn = spawn; V.n_own = ifx_popc64(own); V.nv = 6 * V.n_own + ifx_popc64(spawn);
  return V;
}
IFX_HD ValidSet ifx_valid_set_of_key(const IfxKey& k) {      // from a node key (ifx_key0 / ifx_key1)
  const uint64_t own = k.k0 & IFX_BOARD_MASK, opp = k.k1 & IFX_BOARD_MASK;
  return ifx_valid_set(own, ((k.k1 >> 49) & 1ull) ? (~(own | opp) & IFX_BOARD_MASK) : 0ull);
}
IFX_HD int ifx_rank64(uint64_t m, int cell) { return ifx_popc64(m & ((1ull << cell) - 1ull)); }
// number of valid actions with an index below a (= the compact index of a, if a is valid); 0 <= a <= 343
IFX_HD int ifx_valid_below(const ValidSet& V, int a) {
  const int m = a / IFX_CELLS, cell = a - m * IFX_CELLS;
  if (m < IFX_SPAWN) return m * V.n_own + ifx_rank64(V.own, cell);
  if (m == IFX_SPAWN) return 6 * V.n_own + ifx_rank64(V.spawn, cell);
  return V.nv;
}
// position of the k-th (0-based) set bit of m (k < popc(m)): a six-step binary search on population counts
// (__fns is a software loop on this architecture, ~10x the instructions)
IFX_HD int ifx_nth_set_bit64(uint64_t m, int k) {
  int pos = 0;
  uint32_t w = (uint32_t)m;
  int c = ifx_popc32(w);
  if (k >= c) { k -= c; pos = 32; w = (uint32_t)(m >> 32); }
  c = ifx_popc32(w & 0xFFFFu); if (k >= c) { k -= c; pos += 16; w >>= 16; }
  c = ifx_popc32(w & 0xFFu);   if (k >= c) { k -= c; pos += 8;  w >>= 8; }
  c = ifx_popc32(w & 0xFu);    if (k >= c) { k -= c; pos += 4;  w >>= 4; }
  c = ifx_popc32(w & 0x3u);    if (k >= c) { k -= c; pos += 2;  w >>= 2; }
  c = (int)(w & 1u);           if (k >= c) { pos += 1; }
  return pos;
}
// action index of compact entry ci (0 <= ci < nv)
IFX_HD int ifx_compact_to_action(const ValidSet& V, int ci) {
  const int spread = 6 * V.n_own;
  if (ci < spread) {
    const int m = ci / V.n_own, k = ci - m * V.n_own;
    return m * IFX_CELLS + ifx_nth_set_bit64(V.own, k);
  }
  return IFX_SPAWN * IFX_CELLS + ifx_nth_set_bit64(V.spawn, ci - spread);
}
// per-node header: Ns and the two square roots PUCT needs, refreshed whenever Ns changes (off the select path)
struct alignas(16) IfxHead { int32_t ns; float sq_seen, sq_unseen; int32_t pad; };
IFX_HD IfxHead ifx_head_of(int ns) {
  IfxHead h; h.ns = ns; h.pad = 0;
  // float32(math.sqrt(Ns)) (MCTS.py:122) and float32(math.sqrt(Ns + 1e-8)) (MCTS.py:125).  For every integer
  // 1 <= Ns < 2^17 (the width of a visit count) both equal the correctly rounded float32 square root of Ns
  // (exhaustively checked in tests/test_oracle_mcts.py), so no float64 square root runs on the backup path.
  const float s = ifx_fsqrt((float)ns);
  h.sq_seen = s;
  h.sq_unseen = ns > 0 ? s : 1e-4f;                         // float32(sqrt(1e-8)) = float32(1e-4)
  return h;
}

struct TreeParams {
  float* rows;          // [G][C][3][352]  P | N | Q
  double* q64;          // [G][C][344]
  IfxKey* node_key;     // [G][C]
  IfxHead* node_head;   // [G][C]  Ns + cached square roots
  IfxKey* hash;         // [G][H]
  uint16_t* free_stack; // [G][C]
  IfxGame* games;       // [G]
  IfxU4* path;          // [G][path_stride]  {slot << 9 | row entry, N word, Q bits, Ns} per level (valid while a leaf is
                        // pending); levels beyond the 32 a warp holds in registers live only here
  ifx_state* roots;     // [G]
  ifx_state* leaves;    // [G]
  uint8_t* sym;         // [G][sym_cap][64]  source cell of every destination cell under the random_symmetry of simulation i of
                        // the CURRENT move (filled when the move starts); byte 63 of a row = the symmetry code
  int32_t G, C, H, maxd, path_stride, sym_cap, num_sims, logits, term_budget;
  int32_t vl_mode;            // 1: waves run the leaf-parallel kernel (also for K = 1: a test hook, leaves_per_wave = -1)
  int32_t leaves_per_wave;    // 1: the reference's sequential search; K > 1: leaf-parallel mode (ifx_tree_vl.cuh), K leaf / path areas per game
  float dir_alpha, dir_eps;   // optional Dirichlet root noise (eps = 0: off = the reference's behaviour)
  uint64_t* debug;      // optional [G][8]: ns stamps {start, loads done, expanded, backed up, first level done, end}, depth, probes
  float cpuct;
  uint64_t seed;
};

// 108-bit node key -> table index bits: four 32-bit multiplies and a 32-bit finaliser (the table is indexed by the
// low bits; 64-bit multiplies cost 3-4 instructions each on this datapath)
IFX_HD uint32_t ifx_hash_key(uint64_t k0, uint64_t k1) {
  const uint32_t a = (uint32_t)k0, b = (uint32_t)(k0 >> 32), c = (uint32_t)k1, d = (uint32_t)(k1 >> 32);
  uint32_t h = (a * 0x9E3779B1u) ^ ((b + 0x7F4A7C15u) * 0x85EBCA77u) ^ ((c ^ 0x27D4EB2Fu) * 0xC2B2AE3Du) ^
               ((d + 0x165667B1u) * 0x27D4EB2Fu);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

// Every pool is [G][...]: a warp works on ONE game, so the wave kernel rebases the pointers to its game once and
// passes g = 0 from then on -- the 64-bit `g * stride` products disappear from every address computation.
IFX_HD TreeParams tree_rebase(const TreeParams& P, int g) {
  TreeParams L = P;
  L.rows += (int64_t)g * P.C * 3 * IFX_ROW; L.q64 += (int64_t)g * P.C * 344;
  L.node_key += (int64_t)g * P.C; L.node_head += (int64_t)g * P.C; L.hash += (int64_t)g * P.H;
  L.free_stack += (int64_t)g * P.C; L.games += g; L.path += (int64_t)g * P.path_stride; L.roots += g; L.leaves += g; L.sym += (int64_t)g * P.sym_cap * 64;
  return L;
}

IFX_HD float* tree_row(const TreeParams& P, int g, int slot, int which) {
  return P.rows + (((int64_t)g * P.C + slot) * 3 + which) * IFX_ROW;
}

IFX_HD bool tree_action_bit(uint64_t own, uint64_t spawn, int a) {
  const int m = a / IFX_CELLS, cell = a - m * IFX_CELLS;
  return ((m == IFX_SPAWN ? spawn : own) >> cell) & 1ull;
}

// ---- hash (open addressing, linear probing, never deleted from: rebuilt when nodes are reclaimed) ----
// Entries are probed IFX_PROBE at a time (8 x 16 B = one 128-byte line; at a load factor <= 0.5 the first group
// almost always holds the match or an empty entry).  The first group is requested by tree_hash_issue; the caller
// may do unrelated work before tree_hash_resolve consumes it (and keeps probing if that group decides nothing).
// resolve returns the slot, or -1 with *miss_pos = position of the first empty entry of the probe sequence
// (the insert point).
#define IFX_PROBE 8
IFX_D void tree_hash_issue(const TreeParams& P, int g, uint32_t h, LaneVar<IfxKey>& first) {
  const IfxKey* tab = P.hash + (int64_t)g * P.H;
  IFX_FOR_LANES(l) {
    IfxKey e; e.k0 = 0; e.k1 = 0;                            // lanes beyond the group: neither a match nor empty
    if (l < IFX_PROBE) e = tab[(h + l) & (P.H - 1)];
    first[l] = e;
  }
}
IFX_D int tree_hash_resolve(const TreeParams& P, int g, uint64_t k0, uint64_t k1, uint32_t h, const LaneVar<IfxKey>& first,
                            int* miss_pos) {
  const IfxKey* tab = P.hash + (int64_t)g * P.H;
  for (int probe = 0; probe < P.H; probe += IFX_PROBE) {
    LaneVar<int> match, empty, slot;
    IFX_FOR_LANES(l) {
      IfxKey e = first[l];
      if (probe > 0) { e.k0 = 0; e.k1 = 0; if (l < IFX_PROBE) e = tab[(h + probe + l) & (P.H - 1)]; }
      match[l] = (l < IFX_PROBE && e.k0 == k0 && (e.k1 & IFX_KEY1_MASK) == k1) ? 1 : 0;
      empty[l] = (l < IFX_PROBE && e.k0 == IFX_KEY_EMPTY) ? 1 : 0;
      slot[l] = (int)(e.k1 >> 50);
    }
    const uint32_t mm = lanes_ballot(match), me = lanes_ballot(empty);
    if (mm && (!me || ifx_ffs32(mm) < ifx_ffs32(me))) return lanes_bcast(slot, ifx_ffs32(mm) - 1);
    if (me) { *miss_pos = (int)((h + probe + ifx_ffs32(me) - 1) & (P.H - 1)); return -1; }
  }
  *miss_pos = -1;
  return -1;
}

// ---- the root's rows, mirrored in shared memory for the length of a wave ----
// Every simulation starts at the root, and the root slot is known as soon as the game record has arrived, i.e.
// long before the pending leaf has been expanded and backed up.  Its three rows, header and key are therefore
// copied into shared memory by the TMA engine right away (no registers held, the copy flies under the expansion),
// the backup patches the one edge it changes, and every simulation of the wave selects from the mirror: one
// HBM read of the root per wave and no dependent round trip at the start of a simulation.
struct alignas(16) RootMirror { float rows[3 * IFX_ROW]; IfxHead head; IfxKey key; ifx_state root; uint64_t mbar; uint64_t pad; };

#ifdef IFX_HOSTSIM
static inline void ifx_cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
static inline void ifx_cp_async_wait() {}
#else
IFX_D void ifx_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
IFX_D void ifx_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP in SASS)
IFX_D void ifx_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
                  "r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}
#endif

// Top of the wave: the common fetch, issued by ONE lane as five TMA bulk copies (three rows of ceil(nv/4) 16-byte
// chunks, header, key) that complete on the warp's mbarrier -- six instructions instead of a cp.async loop over all
// lanes.  The source was written by earlier launches, so no proxy fence is needed.
IFX_D void tree_mirror_fetch(const TreeParams& P, int g, int slot, int nv, RootMirror* M) {
  const uint32_t bytes = (uint32_t)((nv + 3) >> 2) * 16u;
  const char* src = reinterpret_cast<const char*>(tree_row(P, g, slot, 0));
#ifdef IFX_HOSTSIM
  for (int which = 0; which < 3; ++which) memcpy(M->rows + which * IFX_ROW, src + which * 4 * IFX_ROW, bytes);
  M->head = P.node_head[(int64_t)g * P.C + slot];
  M->key = P.node_key[(int64_t)g * P.C + slot];
  M->root = P.roots[g];
#else
  if (IFX_LANE0) {
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&M->mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(3u * bytes + 32u + 64u) : "memory");
    if (bytes) {
      ifx_bulk_g2s(M->rows, src, bytes, &M->mbar);
      ifx_bulk_g2s(M->rows + IFX_ROW, src + 4 * IFX_ROW, bytes, &M->mbar);
      ifx_bulk_g2s(M->rows + 2 * IFX_ROW, src + 8 * IFX_ROW, bytes, &M->mbar);
    }
    ifx_bulk_g2s(&M->head, &P.node_head[(int64_t)g * P.C + slot], 16u, &M->mbar);
    ifx_bulk_g2s(&M->key, &P.node_key[(int64_t)g * P.C + slot], 16u, &M->mbar);
    ifx_bulk_g2s(&M->root, &P.roots[g], 64u, &M->mbar);    // the root position every simulation of the wave starts from
  }
#endif
}
IFX_D void tree_mirror_wait(RootMirror* M) {
#ifndef IFX_HOSTSIM
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&M->mbar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(mb) : "memory");
  }
#endif
  IFX_SYNCWARP();
}
// A root expanded in this very wave (once per move at most): its rows were just written with ordinary stores, so
// they are fetched with cp.async (the same proxy) and waited for at once.
IFX_D void tree_mirror_fetch_now(const TreeParams& P, int g, int slot, int nv, RootMirror* M) {
  const int chunks = (nv + 3) >> 2;
  const char* src = reinterpret_cast<const char*>(tree_row(P, g, slot, 0));
  char* dst = reinterpret_cast<char*>(M->rows);
  IFX_FOR_LANES(l) {
    for (int i = l; i < chunks; i += 32) {
      ifx_cp_async16(dst + 16 * i, src + 16 * i);
      ifx_cp_async16(dst + 16 * i + 4 * IFX_ROW, src + 16 * i + 4 * IFX_ROW);
      ifx_cp_async16(dst + 16 * i + 8 * IFX_ROW, src + 16 * i + 8 * IFX_ROW);
    }
    if (l == 0) ifx_cp_async16(&M->head, &P.node_head[(int64_t)g * P.C + slot]);
    if (l == 1) ifx_cp_async16(&M->key, &P.node_key[(int64_t)g * P.C + slot]);
    if (l >= 4 && l < 8) ifx_cp_async16(reinterpret_cast<char*>(&M->root) + 16 * (l - 4), reinterpret_cast<const char*>(&P.roots[g]) + 16 * (l - 4));
  }
  ifx_cp_async_wait();
  IFX_SYNCWARP();
}

// ---- one edge of the backup (MCTS.py:136-142 under NEP-50 promotion) ----
// (raw N, Q32) were loaded ahead of time; q64 is touched only for Python-float edges.  `child` (>= 0) refreshes
// the edge's child hint.  `a` is the edge's compact index in its node's rows.  `mirror` (nullable) receives the
// same N and Q words (the edge belongs to the root).
IFX_HD void tree_update_edge(const TreeParams& P, int g, int slot, int a, uint32_t raw, float q32, double v, bool v_is32,
                             int child, RootMirror* mirror) {
  uint32_t* nrow = reinterpret_cast<uint32_t*>(tree_row(P, g, slot, 1));
  float* qrow = tree_row(P, g, slot, 2);
  double* q64 = P.q64 + ((int64_t)g * P.C + slot) * 344 + a;
  const uint32_t n = raw & IFX_N_COUNT;
  const bool q_is64 = (raw & IFX_N_TAG) != 0;
  const uint32_t hint = child >= 0 ? ((uint32_t)child << IFX_N_HINT_SHIFT) : (raw & (IFX_N_HINT_NONE << IFX_N_HINT_SHIFT));
  bool new_is64;
  float qnew;
  if (n == 0) {                                    // first visit: Q = v, same kind
    new_is64 = !v_is32;
    if (new_is64) *q64 = v;
    qnew = (float)v;
  } else if (q_is64 && !v_is32) {                  // Python float with Python float: float64 throughout
    const double q = ifx_ddiv(ifx_dadd(ifx_dmul((double)n, *q64), v), (double)(n + 1));
    *q64 = q; qnew = (float)q; new_is64 = true;
  } else if (!q_is64) {                            // float32 Q: everything rounds to float32
    qnew = ifx_fdiv(ifx_fadd(ifx_fmul((float)n, q32), (float)v), (float)(n + 1));
    new_is64 = false;
  } else {                                         // Python-float Q meets a float32 value: N*Q in float64, then float32
    qnew = ifx_fdiv(ifx_fadd((float)ifx_dmul((double)n, *q64), (float)v), (float)(n + 1));
    new_is64 = false;
  }
  const uint32_t nnew = (n + 1) | hint | (new_is64 ? IFX_N_TAG : 0u);
  qrow[a] = qnew;
  nrow[a] = nnew;
  if (mirror) {
    mirror->rows[2 * IFX_ROW + a] = qnew;
    reinterpret_cast<uint32_t*>(mirror->rows)[IFX_ROW + a] = nnew;
  }
}

// The path lives one level per lane (lane L holds edge L) as {slot << 9 | compact index, N word, Q bits, Ns}: the
// edge statistics were in registers when the edge was selected and nothing else touches this game's tree
// before its backup, so the backup needs no loads of its own.
// Back `v` (the value returned by the deepest search() call) up the first `depth` edges of the path; levels
// are independent: distinct nodes (the turn strictly grows along a path).  `leaf_slot` is the node the last
// edge leads to (-1: terminal state, no node).  Level 0 is the root: `mirror` (nullable) is kept in step.
IFX_D uint32_t tree_backup_apply(const TreeParams& P, int g, int depth, const LaneVar<IfxU4>& path, double v, bool v_is32,
                                 int leaf_slot, RootMirror* mirror) {
  LaneVar<uint32_t> edge, below;
  LaneVar<int> ovf;
  IFX_FOR_LANES(l) { edge[l] = path[l].x; }
  lanes_shift_down(edge, below);                            // path entry one level deeper
  IFX_FOR_LANES(l) {
    ovf[l] = 0;
    if (l < depth) {
      const int slot = (int)(edge[l] >> 9), a = (int)(edge[l] & 511u);
      const double vl = ((depth - 1 - l) & 1) ? -v : v;     // each level returns -v to its parent (MCTS.py:145)
      const int child = l + 1 < depth ? (int)(below[l] >> 9) : leaf_slot;
      RootMirror* mr = l == 0 ? mirror : nullptr;
      tree_update_edge(P, g, slot, a, path[l].y, ifx_u2f(path[l].z), vl, v_is32, child, mr);
      const IfxHead hd = ifx_head_of((int)path[l].w + 1);   // MCTS.py:144
      P.node_head[(int64_t)g * P.C + slot] = hd;
      if (mr) mr->head = hd;
      ovf[l] = (int)path[l].w + 1 >= IFX_NS_LIMIT ? 1 : 0;  // N(s,a) <= Ns(s): one test guards the 17-bit count field too
    }
  }
  IFX_SYNCWARP();
  return lanes_ballot(ovf);
}

// The same for a path longer than one register chunk (rare: > 32 edges): every chunk is read back from P.path, where
// the descent left it (full chunks are spilled as they fill, the caller spills the last partial one).  Chunk 0 holds
// the root edge (`mirror`).  Kept out of line so that the hot kernel's register allocation does not see it.
IFX_D_NOINLINE uint32_t tree_backup_from_hbm(const TreeParams P, int g, int depth, double v, bool v_is32, int leaf_slot,
                                             RootMirror* mirror) {
  uint32_t ovf = 0;
  IFX_SYNCWARP();
  for (int base = 0; base < depth; base += IFX_PATH_CHUNK) {
    const int n = depth - base < IFX_PATH_CHUNK ? depth - base : IFX_PATH_CHUNK;
    LaneVar<IfxU4> pc;
    IFX_FOR_LANES(l) {
      IfxU4 e = {0u, 0u, 0u, 0u};
      if (l < n) e = P.path[(int64_t)g * P.path_stride + base + l];
      pc[l] = e;
    }
    const int next = base + IFX_PATH_CHUNK < depth ? (int)(P.path[(int64_t)g * P.path_stride + base + IFX_PATH_CHUNK].x >> 9) : leaf_slot;
    const double vc = ((depth - base - n) & 1) ? -v : v;    // sign of level base+l: (depth-1-base-l) = (n-1-l) + (depth-base-n)
    ovf |= tree_backup_apply(P, g, n, pc, vc, v_is32, next, base == 0 ? mirror : nullptr);
  }
  return ovf;
}

// Move type and cell of compact entry ci, found by the whole warp: lane l tests cells l and l+32 of the mask (is it
// set, and is its rank k?), two ballots name the cell -- a dozen instructions with a short dependency chain instead of
// an integer division and a six-step binary search on population counts.
#ifdef IFX_HOSTSIM
static inline float ifx_rcp_approx(float x) { return 1.0f / x; }
#else
IFX_D float ifx_rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
IFX_D void tree_compact_to_move(const ValidSet& V, int ci, int* move_type, int* cell) {
  const int spread = 6 * V.n_own;
  int m = IFX_SPAWN, k = ci - spread;
  uint64_t mask = V.spawn;
  if (ci < spread) {
    // ci / n_own for ci <= 293, n_own <= 49: (ci + 0.5) / n_own is at least 0.5/49 away from an integer, far more
    // than the error of an approximate reciprocal
    m = (int)(((float)ci + 0.5f) * ifx_rcp_approx((float)V.n_own));
    k = ci - m * V.n_own;
    mask = V.own;
  }
  const uint32_t lo = (uint32_t)mask, hi = (uint32_t)(mask >> 32);
  const int plo = ifx_popc32(lo);
  LaneVar<int> hit0, hit1;
  IFX_FOR_LANES(l) {
    const uint32_t below = (1u << l) - 1u;
    hit0[l] = (((lo >> l) & 1u) && ifx_popc32(lo & below) == k) ? 1 : 0;
    hit1[l] = (((hi >> l) & 1u) && plo + ifx_popc32(hi & below) == k) ? 1 : 0;
  }
  const uint32_t b0 = lanes_ballot(hit0), b1 = lanes_ballot(hit1);
  *move_type = m;
  *cell = b0 ? ifx_ffs32(b0) - 1 : 31 + ifx_ffs32(b1);
}

// ---- PUCT arg-max over the valid actions of one node (MCTS.py:114-129) ----
// Group j of a node's rows = compact entries 128j .. 128j+127: lane l owns entries 128j+4l .. +3, fetched as one
// 128-bit load per row.  Group 0 (all a node with <= 128 valid actions has: p95 of random play is 119) is
// requested as soon as the slot is known, together with the header (square roots) and the key; further groups
// cost a round trip each and are rare.
struct RowRegs { LaneVar<IfxF4> p, q; LaneVar<IfxU4> n; IfxHead head; IfxKey key; };

IFX_D void tree_rows_load_group(const float* rows, int j, RowRegs& R) {
  IFX_FOR_LANES(l) {
    const int i = IFX_GROUP * j + 4 * l;
    IfxF4 pv = {0.f, 0.f, 0.f, 0.f}, qv = pv;
    IfxU4 nv4 = {IFX_N_FRESH, IFX_N_FRESH, IFX_N_FRESH, IFX_N_FRESH};
    if (i < IFX_ROW) {
      pv = *reinterpret_cast<const IfxF4*>(rows + i);
      nv4 = *reinterpret_cast<const IfxU4*>(rows + IFX_ROW + i);
      qv = *reinterpret_cast<const IfxF4*>(rows + 2 * IFX_ROW + i);
    }
    R.p[l] = pv; R.n[l] = nv4; R.q[l] = qv;
  }
}
IFX_D void tree_rows_load(const TreeParams& P, int g, int slot, RowRegs& R) {
  tree_rows_load_group(tree_row(P, g, slot, 0), 0, R);
  R.head = P.node_head[(int64_t)g * P.C + slot];
  R.key = P.node_key[(int64_t)g * P.C + slot];
}
IFX_D void tree_rows_from_mirror(const RootMirror* M, RowRegs& R) {
  tree_rows_load_group(M->rows, 0, R);
  R.head = M->head;
  R.key = M->key;
}

// Groups beyond the first (a node with more than 128 valid actions: rare) come from `mirror_rows` (the root mirror) or,
// when that is null, from the node's rows in HBM -- whose address is only worked out if such a group exists.
// Returns the compact index of the chosen edge (or -1), its raw N word (it carries the child hint) and its Q.
IFX_D int tree_select(const TreeParams& P, RowRegs& R, const float* mirror_rows, int g, int slot, int nv, uint32_t* chosen_raw,
                      float* chosen_q) {
  const float sq_seen = R.head.sq_seen, sq_unseen = R.head.sq_unseen;
  LaneVar<float> bu, bq; LaneVar<int> ba; LaneVar<uint32_t> braw;
  IFX_FOR_LANES(l) { bu[l] = -INFINITY; bq[l] = 0.0f; ba[l] = 0x7fffffff; braw[l] = 0u; }
  for (int j = 0; IFX_GROUP * j < nv; ++j) {
    if (j > 0) tree_rows_load_group(mirror_rows ? mirror_rows : tree_row(P, g, slot, 0), j, R);
    IFX_FOR_LANES(l) {
      float best = bu[l], qsel = bq[l]; int arg = ba[l]; uint32_t raw = braw[l];
      // ascending order inside the lane and across groups, strict '>' keeps the lowest index on ties
#define IFX_PUCT(PV, NV, QV, A)                                                                       \
      do { /* straight-line (no divergent branches) so the four independent chains of a lane interleave */ \
        const bool ok_ = (A) < nv;                                                                      \
        const uint32_t n_ = (NV) & IFX_N_COUNT;                                                         \
        const bool seen_ = ok_ && n_ != 0u;                                                             \
        const float cp = ifx_fmul(P.cpuct, (PV));                                                       \
        /* a dummy numerator for edges that do not use the quotient keeps 0/x out of the division's slow path */ \
        const float quo = ifx_fdiv(seen_ ? ifx_fmul(cp, sq_seen) : 1.0f, (float)(1u + n_));             \
        const float u = seen_ ? ifx_fadd((QV), quo) : ifx_fmul(cp, sq_unseen);                          \
        if (ok_ && u > best) { best = u; arg = (A); raw = (NV); qsel = (QV); } } while (0)
      const int base = IFX_GROUP * j + 4 * l;
      IFX_PUCT(R.p[l].x, R.n[l].x, R.q[l].x, base);     IFX_PUCT(R.p[l].y, R.n[l].y, R.q[l].y, base + 1);
      IFX_PUCT(R.p[l].z, R.n[l].z, R.q[l].z, base + 2); IFX_PUCT(R.p[l].w, R.n[l].w, R.q[l].w, base + 3);
#undef IFX_PUCT
      bu[l] = best; ba[l] = arg; braw[l] = raw; bq[l] = qsel;
    }
  }
  float u; int a;
  lanes_argmax(bu, ba, &u, &a);
  if (a == 0x7fffffff) return -1;
  *chosen_raw = lanes_bcast(braw, (a & (IFX_GROUP - 1)) >> 2);   // the lane that owns entry a (in every group)
  *chosen_q = lanes_bcast(bq, (a & (IFX_GROUP - 1)) >> 2);
  return a;
}

// ---- numpy's pairwise float32 sum of a 343-vector held in `buf` ----
// 343 -> [0,80) [80,168) [168,248) [248,343): four blocks x eight strided accumulators = 32 lanes.
IFX_D float tree_pairwise_sum343(const float* buf) {
  LaneVar<float> r;
  IFX_FOR_LANES(l) {
    const int blk = l >> 3, j = l & 7;
    const int base = blk == 0 ? 0 : (blk == 1 ? 80 : (blk == 2 ? 168 : 248));
    const int rows = (blk == 0 || blk == 2) ? 10 : 11;      // 80, 88, 80, 95 (=11 rows + 7 tail)
    float acc = buf[base + j];
    for (int i = 1; i < rows; ++i) acc = ifx_fadd(acc, buf[base + 8 * i + j]);
    r[l] = acc;
  }
  lanes_bfly_fadd(r, 1);            // (r0+r1), (r2+r3), ...
  lanes_bfly_fadd(r, 2);            // (r0+r1)+(r2+r3), ...
  lanes_bfly_fadd(r, 4);            // block sums
  IFX_FOR_LANES(l) {
    if ((l >> 3) == 3) {            // block D: the 95 % 8 = 7 tail elements are added one by one
      float acc = r[l];
      for (int i = 336; i < 343; ++i) acc = ifx_fadd(acc, buf[i]);
      r[l] = acc;
    }
  }
  lanes_bfly_fadd(r, 8);            // A+B, C+D
  lanes_bfly_fadd(r, 16);           // (A+B)+(C+D)
  return lanes_bcast(r, 0);
}

// ---- the net's answer for the pending leaf, staged in shared memory ----
// Per warp: `buf` (352 floats) receives the row the net wrote for this game -- float32 probabilities (exact mode), or
// the binary16 head output (fused mode: 344 halfs in the lower half of buf; the upper half later holds the valid
// logits in compact order) -- and `cells` two 64-byte lists: the node's own cells and its spawn cells in ascending
// order, from which the action of any compact row entry follows in a few instructions.
struct alignas(16) ExpandScratch { float buf[IFX_ROW]; uint8_t own_cell[64]; uint8_t spawn_cell[64]; };

// The row travels in two steps so that its loads can be ISSUED with the first round trip of the wave (before anything
// waits on the game record) and STORED to shared memory once they have landed: a warp issues in order, so a store
// placed right behind its load would hold back every later load by a full round trip.
struct StagedRow { LaneVar<IfxF4> f[3]; };                   // exact mode: 12 floats per lane; fused mode: f[0..1] hold 12 halfs
// exact mode: pi row (float32, >= 343 entries)
IFX_D void tree_stage_policy_load(const float* pi, bool aligned, StagedRow& R) {
  IFX_FOR_LANES(l) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = 4 * l + 128 * j;
      IfxF4 v = {0.f, 0.f, 0.f, 0.f};
      if (aligned) {                                        // padded rows (stride % 4 == 0): one 128-bit load
        if (i < IFX_ROW) v = *reinterpret_cast<const IfxF4*>(pi + i);
      } else {
        if (i < IFX_ACTIONS) v.x = pi[i];
        if (i + 1 < IFX_ACTIONS) v.y = pi[i + 1];
        if (i + 2 < IFX_ACTIONS) v.z = pi[i + 2];
        if (i + 3 < IFX_ACTIONS) v.w = pi[i + 3];
      }
      R.f[j][l] = v;
    }
  }
}
IFX_D void tree_stage_policy_store(const StagedRow& R, float* buf) {
  IFX_FOR_LANES(l) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = 4 * l + 128 * j;
      if (i < IFX_ROW) *reinterpret_cast<IfxF4*>(buf + i) = R.f[j][l];
    }
  }
}
// fused mode: binary16 head row (343 logits + the value pre-activation, 8-byte aligned) -> lower half of buf
IFX_D void tree_stage_head_load(const uint16_t* head, StagedRow& R) {
  IFX_FOR_LANES(l) {
    uint2 h[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = 4 * l + 128 * j;
      h[j].x = 0u; h[j].y = 0u;
      if (i < 344) h[j] = *reinterpret_cast<const uint2*>(head + i);
    }
    IfxF4 a, b;
    a.x = ifx_u2f(h[0].x); a.y = ifx_u2f(h[0].y); a.z = ifx_u2f(h[1].x); a.w = ifx_u2f(h[1].y);
    b.x = ifx_u2f(h[2].x); b.y = ifx_u2f(h[2].y); b.z = 0.0f; b.w = 0.0f;
    R.f[0][l] = a; R.f[1][l] = b;
  }
}
IFX_D void tree_stage_head_store(const StagedRow& R, float* buf) {
  uint16_t* sh = reinterpret_cast<uint16_t*>(buf);
  IFX_FOR_LANES(l) {
    const IfxF4 a = R.f[0][l], b = R.f[1][l];
    uint2 h[3];
    h[0].x = ifx_f2u(a.x); h[0].y = ifx_f2u(a.y); h[1].x = ifx_f2u(a.z); h[1].y = ifx_f2u(a.w);
    h[2].x = ifx_f2u(b.x); h[2].y = ifx_f2u(b.y);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = 4 * l + 128 * j;
      if (i < 344) *reinterpret_cast<uint2*>(sh + i) = h[j];
    }
  }
}

// rank of `cell` among the set bits of a 49-bit mask given as two words
IFX_HD int ifx_rank_lohi(uint32_t lo, uint32_t hi, int plo, int cell) {
  return cell < 32 ? ifx_popc32(lo & ((1u << cell) - 1u)) : plo + ifx_popc32(hi & ((1u << (cell - 32)) - 1u));
}

// action of compact entry ci, from the cell lists (see ExpandScratch)
IFX_D int tree_entry_action(const ExpandScratch* S, int ci, int n_own, int spread, float inv_n_own) {
  if (ci < spread) {
    const int m = (int)(((float)ci + 0.5f) * inv_n_own);    // ci / n_own, see tree_compact_to_move
    return m * IFX_CELLS + (int)S->own_cell[ci - m * n_own];
  }
  return IFX_SPAWN * IFX_CELLS + (int)S->spawn_cell[ci - spread];
}

// ---- expansion of the pending leaf (MCTS.py:89-112) ----
// The work is laid out over the node's VALID actions (86 on average, three per lane), not over the 343 actions of the
// net's row: the masked renormalisation (MCTS.py:95-97) needs an exp / a division / a store only for those.
//   exact mode: policies *= valids happens in place in buf (invalid entries zeroed), policies.sum() is numpy's pairwise
//     float32 sum over the dense vector, then one IEEE division per valid entry -- bit for bit the reference's numbers.
//   fused mode: the reference computes exp(log_softmax(logits)) (NNet.py:94 with InflexionNNet.py:54), masks and
//     renormalises; the softmax denominator cancels in that renormalisation, so e_i = exp(logit_i - max over the valid
//     logits) / sum over the valid e_i is the same number (float32 tolerance; one MUFU.EX2 per prior).
// returns the slot, or -1 with a fault bit set; *leaf_nv = number of valid actions of the new node; *leaf_v (fused
// mode only) = tanh of the value pre-activation
IFX_D int tree_expand(const TreeParams& P, int g, IfxGame& G, const ifx_state& leaf, ExpandScratch* S, const bool fused,
                      int* leaf_nv, float* leaf_v) {
  const ValidSet V = ifx_valid_set(ifx_own(leaf), ifx_spawn_mask(leaf));
  *leaf_nv = V.nv;
  float* buf = S->buf;
  const uint32_t olo = (uint32_t)V.own, ohi = (uint32_t)(V.own >> 32), slo = (uint32_t)V.spawn, shi = (uint32_t)(V.spawn >> 32);
  const int oplo = ifx_popc32(olo), splo = ifx_popc32(slo);
  const int spread = 6 * V.n_own;
  // pop the slot for the NEXT expansion now: the load flies under the whole expansion (its value is used last)
  int next = -1;
  if (G.free_top > 0) next = (int)P.free_stack[(int64_t)g * P.C + (G.free_top - 1)];
  IFX_SYNCWARP();                                           // the staged row (written by other lanes) is complete
  // cell lists, and (exact mode) policies *= valids in place (MCTS.py:95)
  IFX_FOR_LANES(l) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = l + 32 * half;
      if (c < IFX_CELLS) {
        const bool ownb = (V.own >> c) & 1ull, spb = (V.spawn >> c) & 1ull;
        if (ownb) S->own_cell[ifx_rank_lohi(olo, ohi, oplo, c)] = (uint8_t)c;
        if (spb) S->spawn_cell[ifx_rank_lohi(slo, shi, splo, c)] = (uint8_t)c;
        if (!fused) {
          if (!ownb) {
#pragma unroll
            for (int m = 0; m < IFX_SPAWN; ++m) buf[m * IFX_CELLS + c] = 0.0f;
          }
          if (!spb) buf[IFX_SPAWN * IFX_CELLS + c] = 0.0f;
        }
      }
    }
  }
  IFX_SYNCWARP();
  const float inv_n_own = V.n_own > 0 ? ifx_rcp_approx((float)V.n_own) : 0.0f;
  float total, scale = 0.0f, mb = 0.0f;
  const float L2E = 1.4426950408889634f;
  uint16_t* sh = reinterpret_cast<uint16_t*>(buf);          // fused: staged halfs [0,344), compact valid logits [352, 352+nv)
  // A lane's first four entries (all it has when nv <= 128, i.e. nearly always) stay in registers through the three
  // steps -- gather, exp / sum, scale / store -- as four independent chains; entries beyond 128 take a plain loop over
  // the compact copy in shared memory.
  LaneVar<IfxF4> xs;
  if (fused) {
    *leaf_v = ifx_tanh_fast(ifx_h2f(sh[IFX_ACTIONS]));
    LaneVar<float> lmax;
    IFX_FOR_LANES(l) {
      float x[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {                       // (32 t < nv is lane-uniform: with 86 valid actions on average the
        const int ci = l + 32 * t;                          //  fourth pass, often the third, is skipped by the whole warp)
        x[t] = -INFINITY;
        if (32 * t < V.nv && ci < V.nv) x[t] = ifx_h2f(sh[tree_entry_action(S, ci, V.n_own, spread, inv_n_own)]);
      }
      float mx = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
      for (int ci = l + 128; ci < V.nv; ci += 32) {
        const uint16_t hx = sh[tree_entry_action(S, ci, V.n_own, spread, inv_n_own)];
        sh[IFX_ROW + ci] = hx;
        mx = fmaxf(mx, ifx_h2f(hx));
      }
      IfxF4 xv; xv.x = x[0]; xv.y = x[1]; xv.z = x[2]; xv.w = x[3];
      xs[l] = xv; lmax[l] = mx;
    }
    const float m = lanes_max_f32(lmax);
    mb = -m * L2E;
    LaneVar<float> part;
    IFX_FOR_LANES(l) {                                      // exp(x - m) = 2^((x - m) * log2(e)): one multiply-add and one MUFU.EX2
      IfxF4 e = xs[l];                                      // (entries beyond nv hold -inf: 2^-inf = 0)
      e.x = ifx_fast_exp2f(fmaf(e.x, L2E, mb)); e.y = ifx_fast_exp2f(fmaf(e.y, L2E, mb));
      e.z = ifx_fast_exp2f(fmaf(e.z, L2E, mb)); e.w = ifx_fast_exp2f(fmaf(e.w, L2E, mb));
      float acc = (e.x + e.y) + (e.z + e.w);
      for (int ci = l + 128; ci < V.nv; ci += 32) acc += ifx_fast_exp2f(fmaf(ifx_h2f(sh[IFX_ROW + ci]), L2E, mb));
      xs[l] = e; part[l] = acc;
    }
    total = lanes_sum_f32(part);                            // float32 net output: any summation order is within tolerance
  } else {
    total = tree_pairwise_sum343(buf);                      // policies.sum() over the DENSE vector, numpy's order
  }
  const bool uniform = !(total > 0.0f);                     // MCTS.py:98-107: all valid moves masked
  if (uniform) total = (float)V.nv;                         // sum of 0/1 values is exact in any order
  if (fused) scale = ifx_frcp(total);
  const int slot = G.next_slot;
  if (slot < 0) { G.faults |= IFX_F_POOL; return -1; }
  if (G.leaf_hpos < 0) { G.faults |= IFX_F_HASH; return -1; }
  float* prow = tree_row(P, g, slot, 0);
  float* nrow = tree_row(P, g, slot, 1);
  IFX_FOR_LANES(l) {                                        // stores are coalesced: consecutive lanes, consecutive row entries
    if (fused) {
      const IfxF4 e = xs[l];
      if (l < V.nv) prow[l] = uniform ? scale : e.x * scale;
      if (l + 32 < V.nv) prow[l + 32] = uniform ? scale : e.y * scale;
      if (l + 64 < V.nv) prow[l + 64] = uniform ? scale : e.z * scale;
      if (l + 96 < V.nv) prow[l + 96] = uniform ? scale : e.w * scale;
      for (int ci = l + 128; ci < V.nv; ci += 32)
        prow[ci] = uniform ? scale : ifx_fast_exp2f(fmaf(ifx_h2f(sh[IFX_ROW + ci]), L2E, mb)) * scale;
    } else {                                                // policies /= sum (policies = valids in the fallback, MCTS.py:105)
      float v[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int ci = l + 32 * t;
        v[t] = 1.0f;
        if (32 * t < V.nv && ci < V.nv && !uniform) v[t] = buf[tree_entry_action(S, ci, V.n_own, spread, inv_n_own)];
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {                         // (a dummy 1.0 keeps idle lanes out of the division's slow path)
        if (32 * t < V.nv) {                                // lane-uniform: no division for a pass nobody needs
          const float q = ifx_fdiv(v[t], total);
          if (l + 32 * t < V.nv) prow[l + 32 * t] = q;
        }
      }
      for (int ci = l + 128; ci < V.nv; ci += 32)
        prow[ci] = ifx_fdiv(uniform ? 1.0f : buf[tree_entry_action(S, ci, V.n_own, spread, inv_n_own)], total);
    }
    const IfxU4 z = {IFX_N_FRESH, IFX_N_FRESH, IFX_N_FRESH, IFX_N_FRESH};   // Nsa = 0 for every valid action
    for (int c4 = 4 * l; c4 < V.nv; c4 += 128) *reinterpret_cast<IfxU4*>(nrow + c4) = z;
  }
  const uint64_t k0 = ifx_key0(leaf), k1 = ifx_key1(leaf);
  if (IFX_LANE0) {
    IfxKey e; e.k0 = k0; e.k1 = k1 | ((uint64_t)slot << 50);
    P.hash[(int64_t)g * P.H + G.leaf_hpos] = e;             // the position the missed lookup ended on
    IfxKey k; k.k0 = k0; k.k1 = k1;
    P.node_key[(int64_t)g * P.C + slot] = k;
    P.node_head[(int64_t)g * P.C + slot] = ifx_head_of(0);  // Ns[s] = 0 (MCTS.py:110)
  }
  if (G.free_top > 0) G.free_top -= 1;
  G.next_slot = next;
  G.n_live += 1;
  G.n_expand += 1;
  IFX_SYNCWARP();
  return slot;
}

// ---- optional Dirichlet root noise (NOT in the reference; off unless dirichlet_eps > 0) ----
// P(root, .) <- (1 - eps) P + eps * Dir(alpha) over the root's valid actions, once per move, on the device:
// gamma(alpha) variates by Marsaglia-Tsang on alpha + 1 with the U^(1/alpha) boost, every uniform a Philox value
// keyed by (seed; game, move, entry | attempt << 16, IFX_RNG_DIRICHLET), normalised across the warp.
IFX_HD float ifx_u01(uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)
IFX_HD float tree_gamma(uint64_t seed, uint32_t game, uint32_t move, uint32_t entry, float alpha) {
  const float d = (alpha + 1.0f) - (1.0f / 3.0f), c = 1.0f / ifx_fsqrt(9.0f * d);
  float g = d;                                              // (fallback after 8 rejections: the mode region)
  float boost_u = 0.5f;
  for (uint32_t attempt = 0; attempt < 8u; ++attempt) {
    const IfxRand4 r = ifx_philox(seed, game, move, entry | (attempt << 16), IFX_RNG_DIRICHLET);
    if (attempt == 0u) boost_u = ifx_u01(r.w[3]);
    const float x = ifx_fsqrt(-2.0f * ifx_noise_logf(ifx_u01(r.w[0]))) * ifx_cosf(6.2831853071795865f * ifx_u01(r.w[1]));
    const float t = 1.0f + c * x, v = t * t * t;
    if (v > 0.0f && ifx_noise_logf(ifx_u01(r.w[2])) < 0.5f * x * x + d - d * v + d * ifx_noise_logf(v)) { g = d * v; break; }
  }
  return g * ifx_powf(boost_u, 1.0f / alpha);
}
IFX_D_NOINLINE void tree_root_noise_apply(float* prow, uint64_t seed, uint32_t game, uint32_t move, int nv, float alpha, float eps) {
  LaneVar<float> part;
  IFX_FOR_LANES(l) {
    float s = 0.0f;
    for (int i = l; i < nv; i += 32) s += tree_gamma(seed, game, move, (uint32_t)i, alpha);
    part[l] = s;
  }
  const float total = lanes_sum_f32(part);
  if (!(total > 0.0f)) return;
  IFX_FOR_LANES(l) {
    for (int i = l; i < nv; i += 32) {
      const float eta = tree_gamma(seed, game, move, (uint32_t)i, alpha) / total;
      prow[i] = (1.0f - eps) * prow[i] + eps * eta;
    }
  }
  IFX_SYNCWARP();
}

IFX_D void tree_root_noise(const TreeParams& P, int g, const IfxGame& G, int slot, int nv) {
  if (!(P.dir_eps > 0.0f) || slot < 0 || nv <= 0) return;
  tree_root_noise_apply(tree_row(P, g, slot, 0), P.seed, G.game_id, (uint32_t)G.move_idx, nv, P.dir_alpha, P.dir_eps);
}

// random_symmetry of (game, move, sim): k + 6*shift + 42*axis (InflexionGame.py:115-122; MCTS.py:90-92).
// The draw is a pure function of (seed; game, move, sim) -- ~100 lane-uniform instructions of Philox -- and what a leaf
// needs from it is the source cell of each of the 49 destination cells.  Both are computed for a whole move when the
// move starts (tree_begin_move; 32 simulations at a time across the lanes) into a 64-byte row per simulation, so a
// simulation issues two byte loads at an address it knows from its first instruction and nothing depends on them
// until (if) it ends on a new leaf.  Simulations beyond the table (num_sims raised past sym_cap by set_params) draw
// inline.
IFX_HD int tree_symmetry_code(uint64_t seed, uint32_t game, uint32_t move, uint32_t sim) {
  const IfxRand4 r = ifx_philox(seed, game, move, sim, IFX_RNG_SYMMETRY);
  return (int)(r.w[0] % 6u) + 6 * (int)(r.w[1] % 7u) + 42 * (int)(r.w[2] % 3u);
}

// Source cell of every (symmetry, destination cell): 126 x 64 bytes, filled once per process by
// tree_sym_lut_kernel (ifx_mcts_create).
struct SymSrc { LaneVar<int> a, b; };                        // lane l: source cells of destination cells l and l + 32
#ifdef IFX_HOSTSIM
static inline int tree_sym_lut(int sym, int cell) { return cell < IFX_CELLS ? ifx_sym_src(sym, cell) : 0; }
#else
__device__ uint8_t g_ifx_sym_lut[126 * 64];
IFX_D int tree_sym_lut(int sym, int cell) { return (int)g_ifx_sym_lut[sym * 64 + cell]; }
#endif
IFX_D_NOINLINE int tree_symmetry_code_slow(uint64_t seed, uint32_t game, uint32_t move, uint32_t sim) {
  return tree_symmetry_code(seed, game, move, sim);
}
IFX_D void tree_fill_symmetries(const TreeParams& P, int g, uint32_t game, uint32_t move) {
  uint8_t* tab = P.sym + (int64_t)g * P.sym_cap * 64;
  for (int base = 0; base < P.sym_cap; base += 32) {
    LaneVar<int> code;
    IFX_FOR_LANES(l) { code[l] = tree_symmetry_code(P.seed, game, move, (uint32_t)(base + l)); }
    for (int j = 0; j < 32; ++j) {
      const int sym = lanes_bcast(code, j);
      IFX_FOR_LANES(l) {
        uint8_t* row = tab + (int64_t)(base + j) * 64;
        row[l] = (uint8_t)tree_sym_lut(sym, l);
        row[l + 32] = l + 32 == 63 ? (uint8_t)sym : (uint8_t)tree_sym_lut(sym, l + 32);
      }
    }
  }
}

// planes of `st` under the symmetry whose source cells are S -> float32 [4][49] (the net's input row of this game)
IFX_D void tree_write_nn_input(const ifx_state& st, const SymSrc& S, float* out) {
  const uint64_t own = ifx_own(st), opp = ifx_opp(st);
  const float turn = (float)st.turn, spawn = ifx_can_spawn(st) ? 1.0f : 0.0f;
  IFX_FOR_LANES(l) {
    out[l] = (float)((own >> S.a[l]) & 1ull);
    out[IFX_CELLS + l] = (float)((opp >> S.a[l]) & 1ull);
    out[2 * IFX_CELLS + l] = turn;
    out[3 * IFX_CELLS + l] = spawn;
    if (l + 32 < IFX_CELLS) {
      out[l + 32] = (float)((own >> S.b[l]) & 1ull);
      out[IFX_CELLS + l + 32] = (float)((opp >> S.b[l]) & 1ull);
      out[2 * IFX_CELLS + l + 32] = turn;
      out[3 * IFX_CELLS + l + 32] = spawn;
    }
  }
}

// ---- one wave for one game (see ifx_mcts_wave in include/ifx_b200.h) ----
#ifdef IFX_HOSTSIM
static inline uint64_t ifx_now_ns() { return 0; }
#else
IFX_D uint64_t ifx_now_ns() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

template <bool DBG>
IFX_D void tree_wave_game(const TreeParams& P0, int g0, const float* nn_pi, int64_t pi_stride, const float* nn_v,
                          float* nn_in, ExpandScratch* scratch, RootMirror* mirror) {
  const TreeParams P = tree_rebase(P0, g0);
  const int g = 0;
  nn_in += (int64_t)g0 * 196;
  const uint64_t dbg_t0 = DBG ? ifx_now_ns() : 0;
  uint64_t dbg_t1 = 0, dbg_t2 = 0, dbg_t3 = 0, dbg_t4 = 0;
  int dbg_depth = 0, dbg_probes = 0;
  // round trip 1: everything whose address is known up front, issued together (leaf / policy / value are
  // speculative: they are only used when a leaf is pending, which is the common case)
  IfxGame G = P.games[g];
  const ifx_state leaf = P.leaves[g];
  float leaf_v = 0.0f;
  StagedRow staged;
  if (P.logits == 2) {                                      // binary16 head output: exp + tanh are fused into the expansion
    tree_stage_head_load(reinterpret_cast<const uint16_t*>(nn_pi) + (int64_t)g0 * pi_stride, staged);
  } else {
    leaf_v = nn_v[g0];
    const float* pi_row = nn_pi + (int64_t)g0 * pi_stride;
    tree_stage_policy_load(pi_row, (pi_stride & 3) == 0 && pi_stride >= IFX_ROW && ((uintptr_t)nn_pi & 15) == 0, staged);
  }
  LaneVar<IfxU4> path;
  // the first 8 levels of the pending leaf's path travel with the first round trip (one 128-byte line); deeper
  // paths (rare) fetch the rest once the record says so
  IFX_FOR_LANES(l) {
    IfxU4 e = {0u, 0u, 0u, 0u};
    if (l < 8) e = P.path[(int64_t)g * P.path_stride + l];
    path[l] = e;
  }
  // -- everything above is in flight; the first USE of the record comes only now (a warp issues in order) --
  // the root's rows (and the root position) start moving into shared memory the moment the record is here; they land
  // while the pending leaf is expanded and backed up
  int mirror_slot = G.root_slot;
  if (mirror_slot >= 0) tree_mirror_fetch(P, g, mirror_slot, G.root_nv, mirror);
  if (P.logits == 2) tree_stage_head_store(staged, scratch->buf); else tree_stage_policy_store(staged, scratch->buf);
  if (G.pending && G.path_len > 8) {
    IFX_FOR_LANES(l) { if (l >= 8) path[l] = P.path[(int64_t)g * P.path_stride + l]; }
  }
  if (G.faults) { if (mirror_slot >= 0) tree_mirror_wait(mirror); return; }
  if (DBG) dbg_t1 = ifx_now_ns() + (uint64_t)(G.pending & 0) + (uint64_t)(leaf.turn & 0);
  if (G.pending) {
    int leaf_nv = 0;
    const int slot = tree_expand(P, g, G, leaf, scratch, P.logits == 2, &leaf_nv, &leaf_v);
    if (DBG) dbg_t2 = ifx_now_ns();
    if (mirror_slot >= 0) tree_mirror_wait(mirror);         // the backup patches the mirror: it must have landed
    if (slot >= 0) {
      if (G.path_len == 0) {                                // the root itself was the leaf
        G.root_slot = slot; G.root_nv = leaf_nv;
        // optional root noise (off in the reference): the game yields, tree_noise_kernel perturbs the new root's
        // priors before the next wave selects from them
        if (P.dir_eps > 0.0f) G.noise_due = 1;
      }
      // search() of the leaf returns -v, v a float32 (MCTS.py:112; NNet.py:94)
      const uint32_t ovf = G.path_len <= IFX_PATH_CHUNK
          ? tree_backup_apply(P, g, G.path_len, path, -(double)leaf_v, true, slot, mirror_slot >= 0 ? mirror : nullptr)
          : tree_backup_from_hbm(P, g, G.path_len, -(double)leaf_v, true, slot, mirror_slot >= 0 ? mirror : nullptr);
      if (ovf) G.faults |= IFX_F_COUNT;
      G.sims_done += 1; G.n_sims += 1;
    }
    G.pending = 0;
    if (DBG) dbg_t3 = ifx_now_ns();
  } else if (mirror_slot >= 0) {
    tree_mirror_wait(mirror);
  }
  int term_left = P.term_budget > 0 ? P.term_budget : 0x7fffffff;
  while (!G.faults && G.sims_done < P.num_sims && term_left > 0 && !G.noise_due) {
    // the root position: from the mirror (shared memory; it came with the root's rows) -- no round trip at the start
    // of a simulation; straight from the workspace only while the root is not expanded yet (first wave of a move)
    ifx_state st = mirror_slot >= 0 ? mirror->root : P.roots[g];
    int depth = 0;
    // this simulation's random_symmetry (source cells of the lane's two destination cells): requested now, used when
    // (if) the simulation ends on a new leaf
    SymSrc sym_src;
    if (G.sims_done < P.sym_cap) {
      IFX_FOR_LANES(l) { sym_src.a[l] = (int)P.sym[G.sims_done * 64 + l]; sym_src.b[l] = (int)P.sym[G.sims_done * 64 + l + 32]; }
    } else {
      const int sym = tree_symmetry_code_slow(P.seed, G.game_id, (uint32_t)G.move_idx, (uint32_t)G.sims_done);
      IFX_FOR_LANES(l) { sym_src.a[l] = tree_sym_lut(sym, l); sym_src.b[l] = tree_sym_lut(sym, l + 32); }
    }
    int slot = G.root_slot;                                 // -1: the root is not expanded yet
    bool verified = slot >= 0;                              // the cached root slot needs no key check
    RowRegs rows;
    bool inflight = slot >= 0;
    if (inflight) {
      if (mirror_slot != slot) {                            // the root was expanded in this very wave (once per move at most)
        tree_mirror_fetch_now(P, g, slot, G.root_nv, mirror);
        mirror_slot = slot;
      }
      tree_rows_from_mirror(mirror, rows);
    }
    for (;;) {
      if (st.outcome != IFX_ONGOING) {                      // MCTS.py:85-87: terminal, value is a Python number
        uint32_t ovf;
        if (depth <= IFX_PATH_CHUNK) {
          ovf = tree_backup_apply(P, g, depth, path, -ifx_outcome_value(st.outcome), false, -1, mirror_slot >= 0 ? mirror : nullptr);
        } else {                                            // deep path: the last partial chunk joins the others in P.path
          const int base = depth & ~(IFX_PATH_CHUNK - 1);
          IFX_FOR_LANES(l) { if (base + l < depth) P.path[(int64_t)g * P.path_stride + base + l] = path[l]; }
          ovf = tree_backup_from_hbm(P, g, depth, -ifx_outcome_value(st.outcome), false, -1, mirror_slot >= 0 ? mirror : nullptr);
        }
        if (ovf) G.faults |= IFX_F_COUNT;
        G.sims_done += 1; G.n_sims += 1; G.n_term += 1;
        term_left -= 1;
        break;
      }
      const uint64_t k0 = ifx_key0(st), k1 = ifx_key1(st);
      int miss_pos = -1;
      // the rows, header and key of the hinted slot were requested one step ago, in ONE round trip; the key decides
      // whether the hint was right (the reference looks nodes up by planes, not by edge: SURVEY.md H1)
      if (inflight && !verified && !(rows.key.k0 == k0 && rows.key.k1 == k1)) slot = -1;
      if (slot < 0) {
        // Look the state up by key.  Without a verified hint the probe almost always misses (the state is new), so the
        // leaf is written out WHILE the probe is in flight: its planes under this simulation's symmetry to nn_in, the
        // state and the path to the workspace.  If the probe hits after all (a transposition) the descent simply goes
        // on; what was written is overwritten by the leaf this simulation ends on, or never read.
        const uint32_t h = ifx_hash_key(k0, k1);
        LaneVar<IfxKey> first;
        tree_hash_issue(P, g, h, first);
        if (IFX_LANE0) P.leaves[g] = st;
        {                                                   // the chunk in registers (earlier, full chunks are there already)
          const int base = depth & ~(IFX_PATH_CHUNK - 1);
          IFX_FOR_LANES(l) { if (base + l < depth) P.path[(int64_t)g * P.path_stride + base + l] = path[l]; }
        }
        tree_write_nn_input(st, sym_src, nn_in);
        slot = tree_hash_resolve(P, g, k0, k1, h, first, &miss_pos);
        if (DBG) dbg_probes += 1;
        if (slot >= 0) {
          tree_rows_load(P, g, slot, rows);
        } else {                                            // MCTS.py:89: leaf -> wait for the net
          G.path_len = (int16_t)depth; G.pending = 1; G.leaf_hpos = miss_pos;
          break;
        }
      }
      G.n_visits += 1;
      uint32_t raw = 0; float qsel = 0.0f;
      const ValidSet V = ifx_valid_set(ifx_own(st), ifx_spawn_mask(st));
      const bool at_root = depth == 0 && slot == mirror_slot;
      const int ci = tree_select(P, rows, at_root ? mirror->rows : nullptr, g, slot, V.nv, &raw, &qsel);
      if (ci < 0) { G.faults |= IFX_F_NOVALID; break; }
      if (depth >= P.maxd) { G.faults |= IFX_F_DEPTH; break; }
      // request the hinted child's rows FIRST -- the hint came out of the arg-max with the edge -- so that the loads
      // fly under everything that follows: the path record, the entry -> move conversion and the rules step
      const uint32_t edge_word = ((uint32_t)slot << 9) | (uint32_t)ci, node_ns = (uint32_t)rows.head.ns;
      const uint32_t hint = (raw >> IFX_N_HINT_SHIFT) & IFX_N_HINT_NONE;
      slot = (hint != IFX_N_HINT_NONE && (int)hint < P.C) ? (int)hint : -1;
      verified = false;
      inflight = slot >= 0;
      if (inflight) tree_rows_load(P, g, slot, rows);
      IFX_FOR_LANES(l) {
        if (l == (depth & (IFX_PATH_CHUNK - 1))) {
          IfxU4 e; e.x = edge_word; e.y = raw; e.z = ifx_f2u(qsel); e.w = node_ns;
          path[l] = e;
        }
      }
      int mv_type, mv_cell;
      tree_compact_to_move(V, ci, &mv_type, &mv_cell);
      depth += 1;
      if ((depth & (IFX_PATH_CHUNK - 1)) == 0) {            // a full chunk of 32 levels moves to P.path (rare)
        IFX_FOR_LANES(l) { P.path[(int64_t)g * P.path_stride + depth - IFX_PATH_CHUNK + l] = path[l]; }
      }
      if (DBG && depth == 1 && dbg_t4 == 0) dbg_t4 = ifx_now_ns();
      if (DBG) dbg_depth = depth > dbg_depth ? depth : dbg_depth;
      ifx_apply_move_t<true>(st, mv_type, mv_cell);         // MCTS.py:132 (spread line from the constant-memory table)
    }
    if (G.pending) break;
  }
  if (IFX_LANE0) P.games[g] = G;
  if (DBG && IFX_LANE0) {
    uint64_t* d = P.debug + (int64_t)g0 * 8;
    d[0] = dbg_t0; d[1] = dbg_t1; d[2] = dbg_t2; d[3] = dbg_t3; d[4] = dbg_t4; d[5] = ifx_now_ns();
    d[6] = (uint64_t)dbg_depth; d[7] = (uint64_t)dbg_probes;
  }
}

#ifdef IFX_HOSTSIM
static inline uint64_t ifx_cas64(uint64_t* p, uint64_t expect, uint64_t desired) {
  const uint64_t old = *p; if (old == expect) *p = desired; return old;
}
#else
IFX_D uint64_t ifx_cas64(uint64_t* p, uint64_t expect, uint64_t desired) {
  return (uint64_t)atomicCAS(reinterpret_cast<unsigned long long*>(p), (unsigned long long)expect,
                             (unsigned long long)desired);
}
#endif

// ---- start of a move: reclaim nodes the new root can no longer reach, rebuild hash + free stack ----
// A node is reachable only if its turn plane exceeds the root's (the key contains the turn), or it is the
// root itself.  Survivors are re-inserted 32 at a time (one per lane, claimed by CAS).
IFX_D void tree_begin_move(const TreeParams& P, int g, bool clear_all, bool prev_searched = true) {
  IfxGame G = P.games[g];
  const ifx_state root = P.roots[g];
  if (clear_all) {                                          // MCTS.__init__ / reset (MCTS.py:21-31,147-148)
    IfxGame Z; memset(&Z, 0, sizeof(Z));
    Z.game_id = root.tag;
    G = Z;
  } else if (prev_searched) {
    // next getActionProb on the same tables.  The move coordinate of the random draws counts the moves this tree has
    // SEARCHED (the k-th getActionProb call of a reference MCTS object): a root handed over as finished -- Arena gives
    // the search only the games in which its player is to move -- does not advance it.
    G.move_idx += 1;
  }
  IfxKey* tab = P.hash + (int64_t)g * P.H;
  IfxKey* keys = P.node_key + (int64_t)g * P.C;
  uint16_t* fs = P.free_stack + (int64_t)g * P.C;
  IFX_FOR_LANES(l) {
    IfxKey e; e.k0 = IFX_KEY_EMPTY; e.k1 = IFX_KEY_EMPTY;
    for (int i = l; i < P.H; i += 32) tab[i] = e;
  }
  IFX_SYNCWARP();
  const uint64_t r0 = ifx_key0(root), r1 = ifx_key1(root);
  const int rturn = (int)root.turn;
  int free_top = 0, live = 0, root_slot = -1;
  for (int base = 0; base < P.C; base += 32) {
    LaneVar<int> is_live, is_free, is_root;
    IFX_FOR_LANES(l) {
      const int s = base + l;
      int lv = 0, fr = 0, rt = 0;
      uint64_t k0 = IFX_KEY_EMPTY, k1 = 0;
      if (s < P.C) {
        if (!clear_all) { const IfxKey k = keys[s]; k0 = k.k0; k1 = k.k1; }
        if (k0 != IFX_KEY_EMPTY) {
          rt = (k0 == r0 && k1 == r1) ? 1 : 0;
          lv = (ifx_key_turn(k0) > rturn || rt) ? 1 : 0;
        }
        fr = !lv;
        if (fr && (clear_all || k0 != IFX_KEY_EMPTY)) { IfxKey e; e.k0 = IFX_KEY_EMPTY; e.k1 = 0; keys[s] = e; }
        if (lv) {                                           // claim the first empty entry of the probe sequence
          uint32_t idx = ifx_hash_key(k0, k1) & (uint32_t)(P.H - 1);
          for (int tries = 0; tries < P.H; ++tries) {
            if (ifx_cas64(&tab[idx].k0, IFX_KEY_EMPTY, k0) == IFX_KEY_EMPTY) {
              tab[idx].k1 = k1 | ((uint64_t)s << 50);
              break;
            }
            idx = (idx + 1) & (uint32_t)(P.H - 1);
          }
        }
      }
      is_live[l] = lv; is_free[l] = fr; is_root[l] = rt;
    }
    const uint32_t mf = lanes_ballot(is_free), ml = lanes_ballot(is_live), mr = lanes_ballot(is_root);
    IFX_FOR_LANES(l) {
      if (is_free[l]) fs[free_top + ifx_popc32(mf & ((1u << l) - 1u))] = (uint16_t)(base + l);
    }
    free_top += ifx_popc32(mf);
    live += ifx_popc32(ml);
    if (mr) root_slot = base + ifx_ffs32(mr) - 1;
  }
  IFX_SYNCWARP();
  // pop the slot of the next expansion ahead of time
  G.next_slot = -1;
  if (free_top > 0) { G.next_slot = (int)fs[free_top - 1]; free_top -= 1; }
  G.free_top = free_top;
  G.n_live = live;
  G.root_slot = root_slot;
  G.root_nv = ifx_valid_set(ifx_own(root), ifx_spawn_mask(root)).nv;
  G.pending = 0;
  G.path_len = 0;
  G.noise_due = 0;
  G.leaf_hpos = -1;
  G.sims_done = root.outcome == IFX_ONGOING ? 0 : P.num_sims;   // a finished game is never searched
  if (root.outcome == IFX_ONGOING) tree_fill_symmetries(P, g, G.game_id, (uint32_t)G.move_idx);
  if (root.outcome == IFX_ONGOING) tree_root_noise(P, g, G, root_slot, G.root_nv);   // (no-op unless dirichlet_eps > 0)
  if (IFX_LANE0) P.games[g] = G;
}

// ---- root statistics / action choice -----------------------------------------------------------
IFX_D int tree_root_slot(const TreeParams& P, int g) { return P.games[g].root_slot; }

// Action of Coach.py:66-81: temp = int(move+1 < temp_threshold); buf / dbuf hold >= 352 ints / doubles of scratch.
IFX_D int tree_sample_action(const TreeParams& P, int g, int temp_threshold, int* buf, double* dbuf) {
  const IfxGame G = P.games[g];
  const ifx_state root = P.roots[g];
  if (root.outcome != IFX_ONGOING || G.faults) return -1;
  const int slot = G.root_slot;
  if (slot < 0) return -1;
  const uint32_t* nrow = reinterpret_cast<const uint32_t*>(tree_row(P, g, slot, 1));
  const ValidSet V = ifx_valid_set(ifx_own(root), ifx_spawn_mask(root));
  LaneVar<int> lmax, lsum;
  IFX_FOR_LANES(l) {
    int mx = 0, sm = 0;
    for (int a = l; a < IFX_ROW; a += 32) {
      const int c = (a < IFX_ACTIONS && tree_action_bit(V.own, V.spawn, a)) ? (int)(nrow[ifx_valid_below(V, a)] & IFX_N_COUNT) : 0;
      buf[a] = c; mx = c > mx ? c : mx; sm += c;
    }
    lmax[l] = mx; lsum[l] = sm;
  }
  IFX_SYNCWARP();
  const int maxc = lanes_max_i32(lmax), total = lanes_sum_i32(lsum);
  if (total <= 0) return -1;
  const bool temp1 = (G.move_idx + 1) < temp_threshold;
  if (!temp1) {
    // MCTS.py:51-56: uniform choice among the arg-max actions (ascending order)
    int nties = 0;
    for (int base = 0; base < IFX_ROW; base += 32) {
      LaneVar<int> tie;
      IFX_FOR_LANES(l) { tie[l] = (base + l < IFX_ACTIONS && buf[base + l] == maxc) ? 1 : 0; }
      nties += ifx_popc32(lanes_ballot(tie));
    }
    const IfxRand4 r = ifx_philox(P.seed, G.game_id, (uint32_t)G.move_idx, 0u, IFX_RNG_TIEBREAK);
    int pick = (int)(r.w[0] % (uint32_t)nties);
    for (int base = 0; base < IFX_ROW; base += 32) {
      LaneVar<int> tie;
      IFX_FOR_LANES(l) { tie[l] = (base + l < IFX_ACTIONS && buf[base + l] == maxc) ? 1 : 0; }
      uint32_t m = lanes_ballot(tie);
      const int c = ifx_popc32(m);
      if (pick < c) {
        while (pick--) m &= m - 1;
        return base + ifx_ffs32(m) - 1;
      }
      pick -= c;
    }
    return -1;
  }
  // Coach.py:81 np.random.choice(343, p=counts/sum): cdf = cumsum(p); cdf /= cdf[-1]; searchsorted(cdf, u, 'right')
  const IfxRand4 r = ifx_philox(P.seed, G.game_id, (uint32_t)G.move_idx, 0u, IFX_RNG_ACTION);
  const double u = ifx_uniform53(r);
  // The float64 divisions are spread over the lanes (11 each, not 686 on every lane: the FP64 pipe is narrow);
  // only the running sum is sequential, as numpy's cumsum is.
  const double dtot = (double)total;
  IFX_FOR_LANES(l) { for (int a = l; a < IFX_ACTIONS; a += 32) dbuf[a] = ifx_ddiv((double)buf[a], dtot); }   // p = counts / sum
  IFX_SYNCWARP();
  if (IFX_LANE0) {                                           // cdf = cumsum(p)
    double run = 0.0;
    for (int a = 0; a < IFX_ACTIONS; ++a) { run = ifx_dadd(run, dbuf[a]); dbuf[a] = run; }
  }
  IFX_SYNCWARP();
  const double last = dbuf[IFX_ACTIONS - 1];
  LaneVar<int> below;                                        // searchsorted(cdf / cdf[-1], u, 'right') = #{a: cdf[a]/last <= u}
  IFX_FOR_LANES(l) {
    int c = 0;
    for (int a = l; a < IFX_ACTIONS; a += 32) c += (ifx_ddiv(dbuf[a], last) <= u) ? 1 : 0;
    below[l] = c;
  }
  const int idx = lanes_sum_i32(below);
  return idx < IFX_ACTIONS ? idx : IFX_ACTIONS - 1;
}

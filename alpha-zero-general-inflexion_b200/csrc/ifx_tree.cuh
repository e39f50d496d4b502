// Tree kernel core: one warp per game tree.  Restates MCTS.search / getActionProb (MCTS.py:33-145) of the
// reference over structure-of-arrays node pools in HBM.  Everything "uniform" (descent, rules step, keys,
// edge arithmetic) is executed redundantly by the 32 lanes; everything over the 343-action row (PUCT
// arg-max, masked renormalisation, root statistics) is split across lanes and combined with shuffles.
//
// Numerics follow SURVEY.md appendix B exactly (numpy 2.x / NEP-50): float32 PUCT with separate
// round-to-nearest operations, numpy's pairwise float32 sum for the prior renormalisation, and edge
// values that are either float32 (came from the net) or Python floats (came from a terminal state).
#pragma once
#include "ifx_rules.cuh"
#include "ifx_philox.cuh"

struct alignas(16) IfxGame {       // 64 bytes per game
  int32_t sims_done;               // root-level search() calls finished in the current move (MCTS.py:45-46)
  int32_t pending;                 // 1: `leaf` waits for the net
  int32_t faults;                  // IFX_F_* bits
  int32_t move_idx;                // getActionProb calls so far (RNG coordinate)
  int32_t path_len;                // edges on the path to the pending leaf
  int32_t free_top;                // free slots on the stack
  uint32_t game_id;                // RNG stream coordinate
  int32_t n_live;
  uint32_t n_visits, n_expand, n_term, n_sims;   // counters: V, E, terminal leaves, simulations
  int32_t pad[4];
};

struct alignas(16) IfxKey { uint64_t k0, k1; };

struct TreeParams {
  float* rows;          // [G][C][3][352]  P | N | Q
  double* q64;          // [G][C][344]
  IfxKey* node_key;     // [G][C]
  int32_t* node_ns;     // [G][C]
  IfxKey* hash;         // [G][H]
  uint16_t* free_stack; // [G][C]
  IfxGame* games;       // [G]
  uint32_t* path;       // [G][maxd]   slot << 9 | action
  ifx_state* roots;     // [G]
  ifx_state* leaves;    // [G]
  int32_t G, C, H, maxd, num_sims, logits;
  float cpuct;
  uint64_t seed;
};

#define IFX_N_TAG 0x80000000u     // bit 31 of an N entry: Q of this edge is a Python float (kept in q64)

IFX_HD uint32_t ifx_hash_key(uint64_t k0, uint64_t k1) {
  uint64_t h = k0 * 0x9E3779B97F4A7C15ull ^ (k1 + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  return (uint32_t)h;
}

IFX_HD float* tree_row(const TreeParams& P, int g, int slot, int which) {
  return P.rows + (((int64_t)g * P.C + slot) * 3 + which) * IFX_ROW;
}

// ---- hash (open addressing, linear probing, never deleted from: rebuilt when nodes are reclaimed) ----
IFX_D int tree_hash_find(const TreeParams& P, int g, uint64_t k0, uint64_t k1) {
  const IfxKey* tab = P.hash + (int64_t)g * P.H;
  const uint32_t h = ifx_hash_key(k0, k1);
  for (int probe = 0; probe < P.H; probe += 32) {
    LaneVar<int> match, empty, slot;
    IFX_FOR_LANES(l) {
      const IfxKey e = tab[(h + probe + l) & (P.H - 1)];
      match[l] = (e.k0 == k0 && (e.k1 & IFX_KEY1_MASK) == k1) ? 1 : 0;
      empty[l] = e.k0 == IFX_KEY_EMPTY ? 1 : 0;
      slot[l] = (int)(e.k1 >> 50);
    }
    const uint32_t mm = lanes_ballot(match), me = lanes_ballot(empty);
    if (mm && (!me || ifx_ffs32(mm) < ifx_ffs32(me))) return lanes_bcast(slot, ifx_ffs32(mm) - 1);
    if (me) return -1;
  }
  return -1;
}

IFX_D bool tree_hash_insert(const TreeParams& P, int g, uint64_t k0, uint64_t k1, int slot) {
  IfxKey* tab = P.hash + (int64_t)g * P.H;
  const uint32_t h = ifx_hash_key(k0, k1);
  for (int probe = 0; probe < P.H; probe += 32) {
    LaneVar<int> empty;
    IFX_FOR_LANES(l) { empty[l] = tab[(h + probe + l) & (P.H - 1)].k0 == IFX_KEY_EMPTY ? 1 : 0; }
    const uint32_t me = lanes_ballot(empty);
    if (me) {
      if (IFX_LANE0) {
        IfxKey e; e.k0 = k0; e.k1 = k1 | ((uint64_t)slot << 50);
        tab[(h + probe + ifx_ffs32(me) - 1) & (P.H - 1)] = e;
      }
      IFX_SYNCWARP();
      return true;
    }
  }
  return false;
}

// ---- one edge of the backup (MCTS.py:136-142 under NEP-50 promotion) ----
IFX_HD void tree_update_edge(const TreeParams& P, int g, int slot, int a, double v, bool v_is32) {
  uint32_t* nrow = reinterpret_cast<uint32_t*>(tree_row(P, g, slot, 1));
  float* qrow = tree_row(P, g, slot, 2);
  double* q64 = P.q64 + ((int64_t)g * P.C + slot) * 344 + a;
  const uint32_t raw = nrow[a];
  const uint32_t n = raw & ~IFX_N_TAG;
  const bool q_is64 = (raw & IFX_N_TAG) != 0;
  if (n == 0) {                                    // first visit: Q = v, same kind
    if (v_is32) { qrow[a] = (float)v; nrow[a] = 1u; }
    else { *q64 = v; qrow[a] = (float)v; nrow[a] = 1u | IFX_N_TAG; }
  } else if (q_is64 && !v_is32) {                  // Python float with Python float: float64 throughout
    const double q = ifx_ddiv(ifx_dadd(ifx_dmul((double)n, *q64), v), (double)(n + 1));
    *q64 = q; qrow[a] = (float)q; nrow[a] = (n + 1) | IFX_N_TAG;
  } else if (!q_is64) {                            // float32 Q: everything rounds to float32
    const float q = ifx_fdiv(ifx_fadd(ifx_fmul((float)n, qrow[a]), (float)v), (float)(n + 1));
    qrow[a] = q; nrow[a] = n + 1;
  } else {                                         // Python-float Q meets a float32 value: N*Q in float64, then float32
    const float q = ifx_fdiv(ifx_fadd((float)ifx_dmul((double)n, *q64), (float)v), (float)(n + 1));
    qrow[a] = q; nrow[a] = n + 1;
  }
}

// back `v` (the value returned by the deepest search() call) up the first `depth` edges of the path
IFX_D void tree_backup(const TreeParams& P, int g, int depth, double v, bool v_is32) {
  const uint32_t* path = P.path + (int64_t)g * P.maxd;
  IFX_SYNCWARP();
  IFX_FOR_LANES(l) {
    for (int L = l; L < depth; L += 32) {          // levels are independent: distinct nodes (turn strictly grows)
      const uint32_t e = path[L];
      const int slot = (int)(e >> 9), a = (int)(e & 511u);
      const double vl = ((depth - 1 - L) & 1) ? -v : v;     // each level returns -v to its parent (MCTS.py:145)
      tree_update_edge(P, g, slot, a, vl, v_is32);
      P.node_ns[(int64_t)g * P.C + slot] += 1;              // MCTS.py:144
    }
  }
  IFX_SYNCWARP();
}

// ---- PUCT arg-max over the valid actions of one node (MCTS.py:114-129) ----
IFX_D int tree_select(const TreeParams& P, int g, int slot, const ifx_state& st) {
  const float* prow = tree_row(P, g, slot, 0);
  const uint32_t* nrow = reinterpret_cast<const uint32_t*>(tree_row(P, g, slot, 1));
  const float* qrow = tree_row(P, g, slot, 2);
  const int ns = P.node_ns[(int64_t)g * P.C + slot];
  const float sq_seen = (float)ifx_dsqrt((double)ns);
  const float sq_unseen = (float)ifx_dsqrt((double)ns + 1e-8);
  const uint64_t own = ifx_own(st), spawn = ifx_spawn_mask(st);
  LaneVar<float> bu; LaneVar<int> ba;
  IFX_FOR_LANES(l) {
    float best = -INFINITY; int arg = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      const int a = l + 32 * j;
      if (a < IFX_ACTIONS) {
        const int m = a / IFX_CELLS, cell = a - m * IFX_CELLS;
        if (((m == IFX_SPAWN ? spawn : own) >> cell) & 1ull) {
          const float cp = ifx_fmul(P.cpuct, prow[a]);
          const uint32_t n = nrow[a] & ~IFX_N_TAG;
          const float u = n ? ifx_fadd(qrow[a], ifx_fdiv(ifx_fmul(cp, sq_seen), (float)(1u + n)))
                            : ifx_fmul(cp, sq_unseen);
          if (u > best) { best = u; arg = a; }              // ascending a within the lane: strict '>' keeps the lowest
        }
      }
    }
    bu[l] = best; ba[l] = arg;
  }
  float u; int a;
  lanes_argmax(bu, ba, &u, &a);
  return a == 0x7fffffff ? -1 : a;
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

// ---- expansion of the pending leaf (MCTS.py:89-112) ----
// returns the slot, or -1 with a fault bit set
IFX_D int tree_expand(const TreeParams& P, int g, IfxGame& G, const ifx_state& leaf, const float* pi, float* buf) {
  const uint64_t own = ifx_own(leaf), spawn = ifx_spawn_mask(leaf);
  IFX_FOR_LANES(l) {
    for (int a = l; a < IFX_ROW; a += 32) {
      float p = 0.0f;
      if (a < IFX_ACTIONS) {
        const int m = a / IFX_CELLS, cell = a - m * IFX_CELLS;
        if (((m == IFX_SPAWN ? spawn : own) >> cell) & 1ull) p = pi[a];      // policies *= valids
      }
      buf[a] = p;
    }
  }
  IFX_SYNCWARP();
  float total = tree_pairwise_sum343(buf);                  // policies.sum()
  const bool uniform = !(total > 0.0f);                     // MCTS.py:98-107: all valid moves masked
  if (uniform) total = (float)ifx_num_valid(leaf);          // sum of 0/1 values is exact in any order
  if (G.free_top <= 0) { G.faults |= IFX_F_POOL; return -1; }
  const int slot = (int)P.free_stack[(int64_t)g * P.C + (G.free_top - 1)];
  float* prow = tree_row(P, g, slot, 0);
  uint32_t* nrow = reinterpret_cast<uint32_t*>(tree_row(P, g, slot, 1));
  IFX_FOR_LANES(l) {
    for (int a = l; a < IFX_ROW; a += 32) {
      float p = buf[a];
      if (uniform && a < IFX_ACTIONS) {
        const int m = a / IFX_CELLS, cell = a - m * IFX_CELLS;
        p = (((m == IFX_SPAWN ? spawn : own) >> cell) & 1ull) ? 1.0f : 0.0f;
      }
      prow[a] = ifx_fdiv(p, total);                         // policies /= sum
      nrow[a] = 0u;
    }
  }
  const uint64_t k0 = ifx_key0(leaf), k1 = ifx_key1(leaf);
  if (!tree_hash_insert(P, g, k0, k1, slot)) { G.faults |= IFX_F_HASH; return -1; }
  if (IFX_LANE0) {
    IfxKey k; k.k0 = k0; k.k1 = k1;
    P.node_key[(int64_t)g * P.C + slot] = k;
    P.node_ns[(int64_t)g * P.C + slot] = 0;                 // MCTS.py:110
  }
  G.free_top -= 1;
  G.n_live += 1;
  G.n_expand += 1;
  IFX_SYNCWARP();
  return slot;
}

// planes of `st` under the random symmetry drawn for (game, move, sim) -> float32 [4][49] (MCTS.py:90-92)
IFX_D void tree_write_nn_input(const TreeParams& P, const IfxGame& G, const ifx_state& st, float* out) {
  const IfxRand4 r = ifx_philox(P.seed, G.game_id, (uint32_t)G.move_idx, (uint32_t)G.sims_done, IFX_RNG_SYMMETRY);
  const int sym = (int)(r.w[0] % 6u) + 6 * (int)(r.w[1] % 7u) + 42 * (int)(r.w[2] % 3u);
  const uint64_t own = ifx_own(st), opp = ifx_opp(st);
  const float turn = (float)st.turn, spawn = ifx_can_spawn(st) ? 1.0f : 0.0f;
  IFX_FOR_LANES(l) {
    for (int cell = l; cell < IFX_CELLS; cell += 32) {
      const int src = ifx_sym_src(sym, cell);
      out[cell] = (float)((own >> src) & 1ull);
      out[IFX_CELLS + cell] = (float)((opp >> src) & 1ull);
      out[2 * IFX_CELLS + cell] = turn;
      out[3 * IFX_CELLS + cell] = spawn;
    }
  }
}

// ---- one wave for one game (see ifx_mcts_wave in include/ifx_b200.h) ----
IFX_D void tree_wave_game(const TreeParams& P, int g, const float* nn_pi, int64_t pi_stride, const float* nn_v,
                          float* nn_in, float* buf) {
  IfxGame G = P.games[g];
  uint32_t* path = P.path + (int64_t)g * P.maxd;
  if (G.faults) return;
  if (G.pending) {
    const ifx_state leaf = P.leaves[g];
    const int slot = tree_expand(P, g, G, leaf, nn_pi + (int64_t)g * pi_stride, buf);
    if (slot >= 0) {
      // search() of the leaf returns -v, v a float32 (MCTS.py:112; NNet.py:94)
      tree_backup(P, g, G.path_len, -(double)nn_v[g], true);
      G.sims_done += 1; G.n_sims += 1;
    }
    G.pending = 0;
  }
  while (!G.faults && G.sims_done < P.num_sims) {
    ifx_state st = P.roots[g];
    int depth = 0;
    for (;;) {
      if (st.outcome != IFX_ONGOING) {                      // MCTS.py:85-87: terminal, value is a Python number
        tree_backup(P, g, depth, -ifx_outcome_value(st.outcome), false);
        G.sims_done += 1; G.n_sims += 1; G.n_term += 1;
        break;
      }
      const int slot = tree_hash_find(P, g, ifx_key0(st), ifx_key1(st));
      if (slot < 0) {                                       // MCTS.py:89: leaf -> wait for the net
        if (IFX_LANE0) P.leaves[g] = st;
        tree_write_nn_input(P, G, st, nn_in + (int64_t)g * 196);
        G.path_len = depth; G.pending = 1;
        break;
      }
      G.n_visits += 1;
      const int a = tree_select(P, g, slot, st);
      if (a < 0) { G.faults |= IFX_F_NOVALID; break; }
      if (depth >= P.maxd) { G.faults |= IFX_F_DEPTH; break; }
      if (IFX_LANE0) path[depth] = ((uint32_t)slot << 9) | (uint32_t)a;
      depth += 1;
      ifx_apply(st, a);                                     // MCTS.py:132
    }
    if (G.pending) break;
  }
  if (IFX_LANE0) P.games[g] = G;
}

// ---- start of a move: reclaim nodes the new root can no longer reach, rebuild hash + free stack ----
// A node is reachable only if its turn plane exceeds the root's (the key contains the turn), or it is the
// root itself.
IFX_D void tree_begin_move(const TreeParams& P, int g, bool clear_all) {
  IfxGame G = P.games[g];
  const ifx_state root = P.roots[g];
  if (clear_all) {                                          // MCTS.__init__ / reset (MCTS.py:21-31,147-148)
    IfxGame Z; memset(&Z, 0, sizeof(Z));
    Z.game_id = root.tag;
    G = Z;
  } else {
    G.move_idx += 1;                                        // next getActionProb on the same tables
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
  int free_top = 0, live = 0;
  for (int base = 0; base < P.C; base += 32) {
    LaneVar<int> is_live, is_free;
    LaneVar<uint64_t> k0v, k1v;
    IFX_FOR_LANES(l) {
      const int s = base + l;
      int lv = 0, fr = 0;
      uint64_t k0 = IFX_KEY_EMPTY, k1 = 0;
      if (s < P.C) {
        if (!clear_all) { const IfxKey k = keys[s]; k0 = k.k0; k1 = k.k1; }
        if (k0 != IFX_KEY_EMPTY) {
          const int t = ifx_key_turn(k0);
          lv = (t > rturn || (k0 == r0 && k1 == r1)) ? 1 : 0;
        }
        fr = !lv;
        if (fr && (clear_all || k0 != IFX_KEY_EMPTY)) { IfxKey e; e.k0 = IFX_KEY_EMPTY; e.k1 = 0; keys[s] = e; }
      }
      is_live[l] = lv; is_free[l] = fr; k0v[l] = k0; k1v[l] = k1;
    }
    const uint32_t mf = lanes_ballot(is_free);
    uint32_t ml = lanes_ballot(is_live);
    IFX_FOR_LANES(l) {
      if (is_free[l]) fs[free_top + ifx_popc32(mf & ((1u << l) - 1u))] = (uint16_t)(base + l);
    }
    free_top += ifx_popc32(mf);
    live += ifx_popc32(ml);
    while (ml) {                                            // re-insert the survivors
      const int l = ifx_ffs32(ml) - 1;
      ml &= ml - 1;
      tree_hash_insert(P, g, lanes_bcast(k0v, l), lanes_bcast(k1v, l), base + l);
    }
  }
  G.free_top = free_top;
  G.n_live = live;
  G.pending = 0;
  G.path_len = 0;
  G.sims_done = root.outcome == IFX_ONGOING ? 0 : P.num_sims;   // a finished game is never searched
  if (IFX_LANE0) P.games[g] = G;
}

// ---- root statistics / action choice -----------------------------------------------------------
IFX_D int tree_root_slot(const TreeParams& P, int g) {
  const ifx_state root = P.roots[g];
  return tree_hash_find(P, g, ifx_key0(root), ifx_key1(root));
}

// Action of Coach.py:66-81: temp = int(move+1 < temp_threshold); buf holds >= 352 ints of scratch.
IFX_D int tree_sample_action(const TreeParams& P, int g, int temp_threshold, int* buf) {
  const IfxGame G = P.games[g];
  const ifx_state root = P.roots[g];
  if (root.outcome != IFX_ONGOING || G.faults) return -1;
  const int slot = tree_root_slot(P, g);
  if (slot < 0) return -1;
  const uint32_t* nrow = reinterpret_cast<const uint32_t*>(tree_row(P, g, slot, 1));
  LaneVar<int> lmax, lsum;
  IFX_FOR_LANES(l) {
    int mx = 0, sm = 0;
    for (int a = l; a < IFX_ROW; a += 32) {
      const int c = a < IFX_ACTIONS ? (int)(nrow[a] & ~IFX_N_TAG) : 0;
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
  const double dtot = (double)total;
  double last = 0.0;
  for (int a = 0; a < IFX_ACTIONS; ++a) last = ifx_dadd(last, ifx_ddiv((double)buf[a], dtot));
  double run = 0.0;
  int idx = 0;
  for (int a = 0; a < IFX_ACTIONS; ++a) {
    run = ifx_dadd(run, ifx_ddiv((double)buf[a], dtot));
    if (ifx_ddiv(run, last) <= u) idx = a + 1;
  }
  return idx < IFX_ACTIONS ? idx : IFX_ACTIONS - 1;
}

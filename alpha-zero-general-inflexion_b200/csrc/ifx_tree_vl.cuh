// Leaf-parallel search with virtual loss (BASELINE.json north_star: "virtual-loss plus atomic backup"): a PERFORMANCE
// MODE, separate from the reference-exact kernel in ifx_tree.cuh and never used for parity.
//
// The reference runs its simulations strictly one after another and calls the net once per simulation (MCTS.py:45-46,
// 89-93); the exact kernel reproduces that, so a tree contributes ONE row to every net batch.  That is what makes a
// single MCTS object (Arena / pit.py through the object API) or a small set of games net-latency-bound: batch-1
// forwards, numMCTSSims of them per move.  Here a game collects up to K leaves per wave (ifx_mcts_config.
// leaves_per_wave): descent j of a wave sees the edges of the wave's earlier descents as if each had already been
// visited once more and lost (N + 1, W - 1: "virtual loss"), so the K paths spread over the tree instead of repeating
// the same one; the net evaluates G x K rows per call and a move needs about numMCTSSims / K calls.  The virtual visits
// live only in shared memory for the length of a wave -- nothing virtual is ever written to the tables; when the net's
// answers arrive the K leaves are expanded and their values backed up one after another with freshly read edge
// statistics (one warp owns the tree, so the "atomic backup" of a multi-threaded CPU implementation is plain program
// order here).  Visit counts differ from the sequential search by construction (K = 1 reproduces it bit for bit,
// tests/test_virtual_loss.py); root visit distributions stay close (same test).
#pragma once
#include "ifx_tree.cuh"

#define IFX_VL_MAX_LEAVES 16
#define IFX_VL_DEPTH 64            // path levels of an in-flight descent that cast virtual loss (deeper levels cast none)

struct alignas(16) VlScratch {
  uint32_t edge[IFX_VL_MAX_LEAVES][IFX_VL_DEPTH];     // {slot << 9 | row entry} per level of the wave's descents so far
  IfxKey key[IFX_VL_MAX_LEAVES];                      // their leaves' keys (a second descent reaching the same new state stops the wave)
  int32_t len[IFX_VL_MAX_LEAVES];
};

// Back `v` up a path of `depth` edges stored in P.path[path_base ..] (edge words only are trusted: the statistics
// recorded at descent time may be stale, because paths of one wave share edges), 32 levels at a time.
IFX_D uint32_t tree_backup_fresh(const TreeParams& P, int g, int64_t path_base, int depth, double v, bool v_is32, int leaf_slot) {
  uint32_t ovf = 0;
  IFX_SYNCWARP();
  for (int base = 0; base < depth; base += IFX_PATH_CHUNK) {
    const int n = depth - base < IFX_PATH_CHUNK ? depth - base : IFX_PATH_CHUNK;
    LaneVar<IfxU4> pc;
    IFX_FOR_LANES(l) {
      IfxU4 e = {0u, 0u, 0u, 0u};
      if (l < n) {
        e.x = P.path[path_base + base + l].x;
        const int slot = (int)(e.x >> 9), a = (int)(e.x & 511u);
        e.y = reinterpret_cast<const uint32_t*>(tree_row(P, g, slot, 1))[a];
        e.z = ifx_f2u(tree_row(P, g, slot, 2)[a]);
        e.w = (uint32_t)P.node_head[(int64_t)g * P.C + slot].ns;
      }
      pc[l] = e;
    }
    const int next = base + IFX_PATH_CHUNK < depth ? (int)(P.path[path_base + base + IFX_PATH_CHUNK].x >> 9) : leaf_slot;
    const double vc = ((depth - base - n) & 1) ? -v : v;
    ovf |= tree_backup_apply(P, g, n, pc, vc, v_is32, next, nullptr);
    IFX_SYNCWARP();
  }
  return ovf;
}

// PUCT arg-max with virtual loss.  vl[l] = virtual visits of the lane's four entries of group j (packed, 8 bits each),
// vl_node = virtual visits of the node.  Entries without virtual loss use exactly the arithmetic of tree_select.
IFX_D int tree_select_vl(const TreeParams& P, RowRegs& R, const float* rows, int nv, const VlScratch* S, int n_inflight,
                         int depth, int slot, uint32_t* chosen_raw, float* chosen_q) {
  int vl_node = 0;
  for (int i = 0; i < n_inflight; ++i)
    if (depth < S->len[i] && (int)(S->edge[i][depth] >> 9) == slot) vl_node += 1;
  float sq_seen = R.head.sq_seen, sq_unseen = R.head.sq_unseen;
  if (vl_node) {
    const float ns = (float)(R.head.ns + vl_node);
    sq_seen = ifx_fsqrt(ns);
    sq_unseen = sq_seen;
  }
  LaneVar<float> bu, bq; LaneVar<int> ba; LaneVar<uint32_t> braw;
  IFX_FOR_LANES(l) { bu[l] = -INFINITY; bq[l] = 0.0f; ba[l] = 0x7fffffff; braw[l] = 0u; }
  for (int j = 0; IFX_GROUP * j < nv; ++j) {
    if (j > 0) tree_rows_load_group(rows, j, R);
    IFX_FOR_LANES(l) {
      uint32_t vl = 0;                                         // 8 bits per entry of the lane
      if (vl_node) {
        for (int i = 0; i < n_inflight; ++i) {
          if (depth < S->len[i]) {
            const uint32_t e = S->edge[i][depth];
            const int ci = (int)(e & 511u);
            if ((int)(e >> 9) == slot && (ci >> 7) == j && ((ci & 127) >> 2) == l) vl += 1u << (8 * (ci & 3));
          }
        }
      }
      float best = bu[l], qsel = bq[l]; int arg = ba[l]; uint32_t raw = braw[l];
#define IFX_PUCT_VL(PV, NV, QV, A, VL)                                                                  \
      do {                                                                                              \
        const bool ok_ = (A) < nv;                                                                      \
        const uint32_t n_ = (NV) & IFX_N_COUNT, v_ = (VL);                                              \
        const bool seen_ = ok_ && (n_ + v_) != 0u;                                                      \
        const float cp = ifx_fmul(P.cpuct, (PV));                                                       \
        float q_ = (QV);                                                                                \
        if (v_) q_ = ((float)n_ * (n_ ? q_ : 0.0f) - (float)v_) / (float)(n_ + v_);                    \
        const float quo = ifx_fdiv(seen_ ? ifx_fmul(cp, sq_seen) : 1.0f, (float)(1u + n_ + v_));        \
        const float u = seen_ ? ifx_fadd(q_, quo) : ifx_fmul(cp, sq_unseen);                            \
        if (ok_ && u > best) { best = u; arg = (A); raw = (NV); qsel = (QV); } } while (0)
      const int base = IFX_GROUP * j + 4 * l;
      IFX_PUCT_VL(R.p[l].x, R.n[l].x, R.q[l].x, base, vl & 255u);          IFX_PUCT_VL(R.p[l].y, R.n[l].y, R.q[l].y, base + 1, (vl >> 8) & 255u);
      IFX_PUCT_VL(R.p[l].z, R.n[l].z, R.q[l].z, base + 2, (vl >> 16) & 255u); IFX_PUCT_VL(R.p[l].w, R.n[l].w, R.q[l].w, base + 3, vl >> 24);
#undef IFX_PUCT_VL
      bu[l] = best; ba[l] = arg; braw[l] = raw; bq[l] = qsel;
    }
  }
  float u; int a;
  lanes_argmax(bu, ba, &u, &a);
  if (a == 0x7fffffff) return -1;
  *chosen_raw = lanes_bcast(braw, (a & (IFX_GROUP - 1)) >> 2);
  *chosen_q = lanes_bcast(bq, (a & (IFX_GROUP - 1)) >> 2);
  return a;
}

// One wave of the leaf-parallel mode for one game.  K = P.leaves_per_wave.  Row r = g0 * K + j of nn_pi / nn_v / nn_in
// belongs to leaf j of game g0.
IFX_D void tree_wave_game_vl(const TreeParams& P0, int g0, const float* nn_pi, int64_t pi_stride, const float* nn_v,
                             float* nn_in, ExpandScratch* scratch, VlScratch* VS) {
  const int K = P0.leaves_per_wave;
  TreeParams P = tree_rebase(P0, g0);
  P.leaves = P0.leaves + (int64_t)g0 * K;                      // K leaf states / K path areas per game
  P.path = P0.path + (int64_t)g0 * K * P0.path_stride;
  const int g = 0;
  IfxGame G = P.games[g];
  if (G.faults) return;
  const bool fused = P.logits == 2;
  // ---- phase A: the net has answered the pending leaves: expand each, back its value up with fresh statistics ----
  const int n_pending = G.pending;
  for (int j = 0; j < n_pending; ++j) {
    const ifx_state leaf = P.leaves[j];
    const int64_t row = (int64_t)g0 * K + j;
    StagedRow staged;
    float leaf_v = 0.0f;
    if (fused) {
      tree_stage_head_load(reinterpret_cast<const uint16_t*>(nn_pi) + row * pi_stride, staged);
    } else {
      leaf_v = nn_v[row];
      tree_stage_policy_load(nn_pi + row * pi_stride, (pi_stride & 3) == 0 && pi_stride >= IFX_ROW && ((uintptr_t)nn_pi & 15) == 0, staged);
    }
    // the leaf may have been expanded meanwhile (by an earlier leaf of this wave reaching the same state through another
    // path): look it up again; the position of the first empty entry is where it goes
    const uint64_t k0 = ifx_key0(leaf), k1 = ifx_key1(leaf);
    const uint32_t h = ifx_hash_key(k0, k1);
    LaneVar<IfxKey> first;
    tree_hash_issue(P, g, h, first);
    IFX_SYNCWARP();
    if (fused) tree_stage_head_store(staged, scratch->buf); else tree_stage_policy_store(staged, scratch->buf);
    int miss_pos = -1;
    const int found = tree_hash_resolve(P, g, k0, k1, h, first, &miss_pos);
    if (found >= 0) continue;                                 // (that evaluation is dropped; the simulation is not counted)
    G.leaf_hpos = miss_pos;
    int leaf_nv = 0;
    const int path_len = (int)leaf.reserved[0];
    const int slot = tree_expand(P, g, G, leaf, scratch, fused, &leaf_nv, &leaf_v);
    if (slot < 0) break;
    if (path_len == 0) { G.root_slot = slot; G.root_nv = leaf_nv; }
    if (tree_backup_fresh(P, g, (int64_t)j * P.path_stride, path_len, -(double)leaf_v, true, slot)) G.faults |= IFX_F_COUNT;
    G.sims_done += 1; G.n_sims += 1;
  }
  G.pending = 0;
  // ---- phase B: collect up to K new leaves; later descents see the earlier ones' edges with a virtual loss ----
  int n_leaf = 0;
  // terminal simulations need no net: a wave may run K times the sequential mode's budget of them, so that a game whose
  // search keeps ending in terminal states (a forced win at the root) still completes about K simulations per wave
  int term_left = P.term_budget > 0 ? P.term_budget * K : 0x7fffffff;
  const ifx_state root = P.roots[g];
  while (!G.faults && n_leaf < K && G.sims_done + n_leaf < P.num_sims && term_left > 0) {
    ifx_state st = root;
    int depth = 0, slot = G.root_slot;
    bool verified = slot >= 0, inflight = slot >= 0, stop = false;
    const int sim = G.sims_done + n_leaf;                      // index of this simulation within the move (RNG coordinate)
    SymSrc sym_src;
    if (sim < P.sym_cap) {
      IFX_FOR_LANES(l) { sym_src.a[l] = (int)P.sym[sim * 64 + l]; sym_src.b[l] = (int)P.sym[sim * 64 + l + 32]; }
    } else {
      const int sym = tree_symmetry_code_slow(P.seed, G.game_id, (uint32_t)G.move_idx, (uint32_t)sim);
      IFX_FOR_LANES(l) { sym_src.a[l] = tree_sym_lut(sym, l); sym_src.b[l] = tree_sym_lut(sym, l + 32); }
    }
    RowRegs rows;
    if (inflight) tree_rows_load(P, g, slot, rows);
    const int64_t path_base = (int64_t)n_leaf * P.path_stride;  // this descent's path area (the next free leaf slot)
    for (;;) {
      if (st.outcome != IFX_ONGOING) {                        // terminal: backed up at once, no net needed
        if (tree_backup_fresh(P, g, path_base, depth, -ifx_outcome_value(st.outcome), false, -1)) G.faults |= IFX_F_COUNT;
        G.sims_done += 1; G.n_sims += 1; G.n_term += 1;
        term_left -= 1;
        break;
      }
      const uint64_t k0 = ifx_key0(st), k1 = ifx_key1(st);
      int miss_pos = -1;
      if (inflight && !verified && !(rows.key.k0 == k0 && rows.key.k1 == k1)) slot = -1;
      if (slot < 0) {
        const uint32_t h = ifx_hash_key(k0, k1);
        LaneVar<IfxKey> first;
        tree_hash_issue(P, g, h, first);
        slot = tree_hash_resolve(P, g, k0, k1, h, first, &miss_pos);
        if (slot >= 0) {
          tree_rows_load(P, g, slot, rows);
        } else {
          // a new state.  If an earlier descent of this wave already waits for the net on the very same state, this
          // descent is abandoned and the wave stops collecting (the state will be a node next wave)
          for (int i = 0; i < n_leaf; ++i) if (VS->key[i].k0 == k0 && VS->key[i].k1 == k1) stop = true;
          if (!stop) {
            st.reserved[0] = (uint32_t)depth;
            if (IFX_LANE0) { P.leaves[n_leaf] = st; VS->key[n_leaf].k0 = k0; VS->key[n_leaf].k1 = k1; VS->len[n_leaf] = depth < IFX_VL_DEPTH ? depth : IFX_VL_DEPTH; }
            tree_write_nn_input(st, sym_src, nn_in + ((int64_t)g0 * K + n_leaf) * 196);
            IFX_FOR_LANES(l) {                                 // its edges cast virtual loss on the descents that follow
              for (int d = l; d < depth && d < IFX_VL_DEPTH; d += 32) VS->edge[n_leaf][d] = P.path[path_base + d].x;
            }
            IFX_SYNCWARP();
            n_leaf += 1;
          }
          break;
        }
      }
      G.n_visits += 1;
      uint32_t raw = 0; float qsel = 0.0f;
      const ValidSet V = ifx_valid_set(ifx_own(st), ifx_spawn_mask(st));
      const int ci = tree_select_vl(P, rows, tree_row(P, g, slot, 0), V.nv, VS, n_leaf, depth, slot, &raw, &qsel);
      if (ci < 0) { G.faults |= IFX_F_NOVALID; break; }
      if (depth >= P.maxd) { G.faults |= IFX_F_DEPTH; break; }
      if (IFX_LANE0) {                                         // the path goes straight to its area in the workspace
        IfxU4 e; e.x = ((uint32_t)slot << 9) | (uint32_t)ci; e.y = raw; e.z = ifx_f2u(qsel); e.w = (uint32_t)rows.head.ns;
        P.path[path_base + depth] = e;
      }
      IFX_SYNCWARP();
      const uint32_t hint = (raw >> IFX_N_HINT_SHIFT) & IFX_N_HINT_NONE;
      slot = (hint != IFX_N_HINT_NONE && (int)hint < P.C) ? (int)hint : -1;
      verified = false;
      inflight = slot >= 0;
      if (inflight) tree_rows_load(P, g, slot, rows);
      int mv_type, mv_cell;
      tree_compact_to_move(V, ci, &mv_type, &mv_cell);
      depth += 1;
      ifx_apply_move_t<true>(st, mv_type, mv_cell);
    }
    if (stop) break;
  }
  G.pending = n_leaf;
  if (IFX_LANE0) P.games[g] = G;
}

// Bitboard Inflexion rules: device functions shared by the rules kernels and the tree kernel.
// Semantics follow inflexion/InflexionGame.py of the reference (cited per function); the representation
// (five 49-bit planes) is new.
#pragma once
#include "ifx_common.cuh"

// (dr, dq) of move number m, InflexionGame.py:15-21
IFX_HD int ifx_dir_dr(int m) { return m == 0 || m == 4 ? 1 : (m == 1 || m == 5 ? -1 : 0); }
IFX_HD int ifx_dir_dq(int m) { return m == 2 || m == 5 ? 1 : (m == 3 || m == 4 ? -1 : 0); }

IFX_HD uint64_t ifx_own(const ifx_state& s) { return s.player ? s.blue : s.red; }
IFX_HD uint64_t ifx_opp(const ifx_state& s) { return s.player ? s.red : s.blue; }
IFX_HD uint64_t ifx_empty(const ifx_state& s) { return ~(s.red | s.blue) & IFX_BOARD_MASK; }
IFX_HD int ifx_total_power(const ifx_state& s) {            // InflexionGame.py:347-348
  return ifx_popc64(s.p0) + 2 * ifx_popc64(s.p1) + 4 * ifx_popc64(s.p2);
}
IFX_HD bool ifx_can_spawn(const ifx_state& s) { return s.total_power <= IFX_MAX_POWER_AT_SPAWN; }
IFX_HD int ifx_cell_power(const ifx_state& s, int cell) {
  return (int)((s.p0 >> cell) & 1) + 2 * (int)((s.p1 >> cell) & 1) + 4 * (int)((s.p2 >> cell) & 1);
}

// valid_actions_mask, InflexionGame.py:93-100: spreads = own stacks (all six directions always legal),
// spawn = empty cells iff total power <= 48.
IFX_HD uint64_t ifx_spawn_mask(const ifx_state& s) { return ifx_can_spawn(s) ? ifx_empty(s) : 0ull; }
IFX_HD bool ifx_action_valid(const ifx_state& s, int action) {
  if (action < 0 || action >= IFX_ACTIONS) return false;
  const int m = action / IFX_CELLS, cell = action - m * IFX_CELLS;
  const uint64_t mask = m == IFX_SPAWN ? ifx_spawn_mask(s) : ifx_own(s);
  return (mask >> cell) & 1ull;
}
IFX_HD int ifx_num_valid(const ifx_state& s) { return 6 * ifx_popc64(ifx_own(s)) + ifx_popc64(ifx_spawn_mask(s)); }

// execute_move, InflexionGame.py:273-310 (the action must be valid).  Updates s in place:
// board, outcome (relative to the NEW side to move), turn+1, player flipped.
// the k cells a stack of power k at `cell` lands on when spread in direction m (k <= 6 < 7: all distinct, none the source)
IFX_HD uint64_t ifx_spread_line(int m, int cell, int k) {
  const int dr = ifx_dir_dr(m), dq = ifx_dir_dq(m);
  int r = cell / IFX_N, q = cell - r * IFX_N;
  uint64_t line = 0;
  for (int i = 0; i < k; ++i) {
    r += dr; q += dq;
    r = r < 0 ? r + IFX_N : (r >= IFX_N ? r - IFX_N : r);
    q = q < 0 ? q + IFX_N : (q >= IFX_N ? q - IFX_N : q);
    line |= 1ull << (r * IFX_N + q);
  }
  return line;
}
#if !defined(IFX_HOSTSIM) && defined(IFX_WITH_LINE_LUT)
// The same masks as a table in constant memory, [6 directions][49 cells][6 powers] (14 KB, filled by ifx_mcts_create):
// ONE uniform-indexed constant load where every lane of a warp steps the same board (the tree kernels); the rules
// kernels, one state per thread, keep the loop (divergent constant addresses would serialise).
__constant__ uint64_t g_ifx_line_lut[6 * IFX_CELLS * 6];
#endif

template <bool LUT>
IFX_HD void ifx_apply_move_t(ifx_state& s, int m, int cell) {   // m = move type 0..6, cell = r*7+q
  const uint64_t src = 1ull << cell;
  uint64_t own = ifx_own(s), opp = ifx_opp(s);
  uint64_t p0 = s.p0, p1 = s.p1, p2 = s.p2;
  int total = (int)s.total_power;                           // kept incrementally (three 64-bit popcounts otherwise)
  if (m == IFX_SPAWN) {                                      // :278-280  board[r,q] = player
    own |= src; p0 |= src;
    total += 1;
  } else {                                                   // :281-289
    const int k = ifx_cell_power(s, cell);
    uint64_t line;                                           // the k cells the stack lands on
#if !defined(IFX_HOSTSIM) && defined(IFX_WITH_LINE_LUT)
    if (LUT) line = k > 0 ? g_ifx_line_lut[(m * IFX_CELLS + cell) * 6 + (k - 1)] : 0ull;
    else
#endif
      line = ifx_spread_line(m, cell, k);
    // |x| + 1 on the three bit planes, 7 -> removed (:287-288)
    const uint64_t c0 = p0 & line;  p0 ^= line;
    const uint64_t c1 = p1 & c0;    p1 ^= c0;
    p2 ^= c1;
    const uint64_t over = p0 & p1 & p2;                      // reached 7: stack is destroyed
    total -= 7 * ifx_popc64(over);                           // a spread moves the stack's k units onto k cells: the sum only
                                                             // changes where a cell reaches 7 and vanishes
    const uint64_t gone = over | src;                        // :289 source cell emptied
    p0 &= ~gone; p1 &= ~gone; p2 &= ~gone;
    own = (own | line) & ~gone;
    opp &= ~(line | gone);
  }
  s.p0 = p0; s.p1 = p1; s.p2 = p2;
  if (s.player) { s.blue = own; s.red = opp; } else { s.red = own; s.blue = opp; }
  s.total_power = (uint8_t)total;
  // outcome for the mover, first match wins (:294-306)
  int out = IFX_ONGOING;
  if (m != IFX_SPAWN && opp == 0) {
    out = IFX_WON;
  } else if (s.turn >= s.max_turns) {
    // power_diff(player) = player * sum(board) (:312-317) = own power - opponent power
    const int own_pw = ifx_popc64(p0 & own) + 2 * ifx_popc64(p1 & own) + 4 * ifx_popc64(p2 & own);
    const int diff = 2 * own_pw - (int)s.total_power;
    out = diff >= 2 ? IFX_WON : (diff <= -2 ? IFX_LOST : IFX_DRAW);
  } else if ((own | opp) == 0) {
    out = IFX_DRAW;
  }
  s.turn = (uint16_t)(s.turn + 1);
  s.player ^= 1;                                             // :310; the setter flips WON <-> LOST (Game.py:59-62)
  s.outcome = (uint8_t)(out == IFX_WON ? IFX_LOST : (out == IFX_LOST ? IFX_WON : out));
}

IFX_HD void ifx_apply_move(ifx_state& s, int m, int cell) { ifx_apply_move_t<false>(s, m, cell); }
IFX_HD void ifx_apply(ifx_state& s, int action) {
  const int m = action / IFX_CELLS;
  ifx_apply_move_t<false>(s, m, action - m * IFX_CELLS);
}

// Value of a terminal state as MCTS.py:87 reads it (GameOutcome.value, flags.py:32-36).
IFX_HD double ifx_outcome_value(int outcome) {
  return outcome == IFX_WON ? 1.0 : (outcome == IFX_LOST ? -1.0 : (outcome == IFX_DRAW ? 1e-4 : 0.0));
}

// Source cell of board symmetry `sym` = k + 6*shift + 42*axis for destination cell (r,q):
// out[r,q] = in[src] where out = translate(rotate(in, k), shift, axis)
// (InflexionGame.py:115-122, 124-196; index maps verified in SURVEY.md appendix C).
// x mod 7 for -14 <= x <= 34 (every caller below): three integer operations instead of a signed remainder
IFX_HD int ifx_mod7(int x) { x += 14; return x - IFX_N * ((x * 37) >> 8); }
IFX_HD int ifx_rot_src(int k, int r, int q) {
  const int s = ifx_mod7(r + q);
  int rr, qq;
  switch (k) {
    case 0: rr = r; qq = q; break;
    case 1: rr = -s; qq = r; break;
    case 2: rr = -q; qq = s; break;
    case 3: rr = -r; qq = -q; break;
    case 4: rr = s; qq = -r; break;
    default: rr = q; qq = -s; break;
  }
  return ifx_mod7(rr) * IFX_N + ifx_mod7(qq);
}
IFX_HD int ifx_sym_src(int sym, int cell) {
  const int k = sym % 6, shift = (sym / 6) % IFX_N, axis = sym / 42;
  int r = cell / IFX_N, q = cell - r * IFX_N;
  // translate: np.roll(x, t, axis): out[i] = in[i - t]; axis 's' rolls q by +t and r by -t (:186-195)
  if (axis == 0) r = ifx_mod7(r - shift);
  else if (axis == 1) q = ifx_mod7(q - shift);
  else { r = ifx_mod7(r + shift); q = ifx_mod7(q - shift); }
  return ifx_rot_src(k, r, q);
}
// the i-th of the 36 forms of symmetries() (:102-113): 0 identity, 1..5 rotations k=i, then for k=1..5,
// j=1..6: translate(rotate(k), j, 'r').
IFX_HD int ifx_sym36(int i) {
  if (i < 6) return i;
  const int k = (i - 6) / 6 + 1, j = (i - 6) % 6 + 1;
  return k + 6 * j;
}

// Node key = to_planes().tobytes() (MCTS.py:48,83): own mask, opponent mask, turn, can_spawn -- 108 bits.
// k0 = own | turn << 49;  k1 = opp | can_spawn << 49   (bits 50..63 of k1 are free: the hash stores the slot there)
IFX_HD uint64_t ifx_key0(const ifx_state& s) { return ifx_own(s) | ((uint64_t)s.turn << 49); }
IFX_HD uint64_t ifx_key1(const ifx_state& s) { return ifx_opp(s) | ((uint64_t)(ifx_can_spawn(s) ? 1 : 0) << 49); }
#define IFX_KEY1_MASK ((1ull << 50) - 1ull)
#define IFX_KEY_EMPTY (~0ull)
IFX_HD int ifx_key_turn(uint64_t k0) { return (int)(k0 >> 49); }

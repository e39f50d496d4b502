// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11).
// Counter-based generator for every random draw on the path; the counter is (game, move, sim, purpose)
// so a draw does not depend on the order in which games are processed.
#pragma once
#include "ifx_common.cuh"

#define IFX_RNG_SYMMETRY 0u   // InflexionGame.py:120-121: k = w0 % 6, shift = w1 % 7, axis = w2 % 3
#define IFX_RNG_TIEBREAK 1u   // MCTS.py:53: index = w0 % number_of_ties
#define IFX_RNG_ACTION 2u     // Coach.py:81: u = ((w0>>5)*2^26 + (w1>>6)) / 2^53

//      purpose 3u            // RandomPlayer (csrc/ifx_rules_api.cu)
#define IFX_RNG_DIRICHLET 4u  // optional root noise (not in the reference): gamma variates, see tree_root_noise

struct IfxRand4 { uint32_t w[4]; };

IFX_HD IfxRand4 ifx_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  IfxRand4 out; out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
  return out;
}

IFX_HD double ifx_uniform53(const IfxRand4& r) {
  return ((double)(r.w[0] >> 5) * 67108864.0 + (double)(r.w[1] >> 6)) / 9007199254740992.0;
}

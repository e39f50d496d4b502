/* ifx_b200.h -- C ABI of libifx_b200.so: the Inflexion rules engine and batched MCTS on sm_100a.
 *
 * The reference (Meatssauce/alpha-zero-general-inflexion) is pure Python and has NO FFI boundary for this
 * path; the boundary is designed here and each entry point cites the Python interface it replaces
 * (paths relative to the reference root).  The Python classes in alpha-zero-general-inflexion_b200/
 * (InflexionGame, MCTS, ...) bind these symbols with ctypes and keep the reference's API surface.
 *
 * Conventions
 *   - plain C, no torch types; pointers are DEVICE pointers unless the function name ends in _host;
 *   - every device entry point is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default);
 *   - return value: 0 = ok, negative = IFX_E_* (argument / launch error).  Per-game runtime faults
 *     (pool exhausted, no valid action, ...) are reported through ifx_mcts_status, never by aborting;
 *   - the caller owns every buffer (including the MCTS workspace); the library frees nothing it did not
 *     allocate.  One ifx_mcts handle per GPU/process; a handle is not thread-safe.
 */
#ifndef IFX_B200_H
#define IFX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IFX_API __attribute__((visibility("default")))

/* ---- error codes --------------------------------------------------------------------------- */
#define IFX_OK 0
#define IFX_E_ARG (-1)      /* bad argument (null pointer, size, alignment)                      */
#define IFX_E_CUDA (-2)     /* CUDA runtime error; see ifx_last_cuda_error()                     */
#define IFX_E_NODEVICE (-3) /* no CUDA device                                                    */

/* ---- per-game fault bits (ifx_mcts_status column 2) ----------------------------------------- */
#define IFX_F_POOL 1     /* node pool exhausted: raise nodes_per_game                            */
#define IFX_F_NOVALID 2  /* side to move has no valid action (the reference crashes, MCTS.py:116,131) */
#define IFX_F_DEPTH 4    /* search path deeper than max_depth                                    */
#define IFX_F_HASH 8     /* hash table full                                                      */
#define IFX_F_BADACTION 16 /* advance/step called with an action that is not valid               */
#define IFX_F_COUNT 32   /* a node's visit count reached 2^17 (the width of the count field)     */

/* ---- outcome codes, relative to the side to move (flags.py:32-44, Game.py:64-72) ------------- */
#define IFX_ONGOING 0
#define IFX_DRAW 1  /* value 1e-4 */
#define IFX_WON 2   /* value +1   */
#define IFX_LOST 3  /* value -1   */

/* Packed game state, 64 bytes, 16-byte aligned.  Replaces the attributes of InflexionGame
 * (inflexion/InflexionGame.py:57-69, Game.py:21-25): _board int64[7,7] in [-6,6] (sign = owner),
 * _player, _curr_turn, _outcome, _max_turns.  Bit i = r*7+q of a mask is cell (r,q).
 * Stack power of a cell = p0 + 2*p1 + 4*p2 (1..6); red/blue are the ownership masks. */
typedef struct ifx_state {
  uint64_t red, blue;      /* cells owned by RED (+) / BLUE (-)                                  */
  uint64_t p0, p1, p2;     /* bit planes of |power|                                              */
  uint16_t turn;           /* _curr_turn                                                         */
  uint16_t max_turns;      /* _max_turns                                                         */
  uint8_t player;          /* side to move: 0 = RED, 1 = BLUE                                    */
  uint8_t outcome;         /* IFX_ONGOING/DRAW/WON/LOST w.r.t. the side to move                  */
  uint8_t total_power;     /* sum |board| (cached; InflexionGame.py:347)                         */
  uint8_t flags;           /* reserved                                                           */
  uint32_t tag;            /* caller tag; for MCTS roots: the game id (RNG stream coordinate)    */
  uint32_t reserved[3];
} ifx_state;

/* ---- library ------------------------------------------------------------------------------- */
IFX_API int ifx_version(void);
IFX_API const char* ifx_last_cuda_error(void);
IFX_API int64_t ifx_launch_count(void); /* kernels launched by this library so far (bench evidence) */
IFX_API int ifx_is_hostsim(void);       /* 1 only in the GPU-less logic-check build used by tests/   */

/* ---- rules engine: one thread per state ------------------------------------------------------ */

/* valid_actions_mask (InflexionGame.py:93-100) fused with to_next_state/execute_move
 * (InflexionGame.py:75-82, 273-310).  masks[i][m] (m = 0..6) is the 49-bit cell mask of move type m for
 * states_in[i] (bit r*7+q  <=>  action m*49+r*7+q is valid).  states_out[i] = states_in[i] after
 * actions[i]; status[i] = 0, or IFX_F_BADACTION (state copied through unchanged) when the action is not
 * valid, which is where the reference raises ValueError / asserts (InflexionGame.py:279,282,291).
 * actions may be NULL (mask only) and masks may be NULL (step only). */
IFX_API int ifx_rules_mask_step(const ifx_state* states_in, const int32_t* actions, int64_t n,
                                uint64_t* masks /* [n][7] */, ifx_state* states_out, int32_t* status,
                                void* stream);

/* to_planes (InflexionGame.py:84-91) with an optional board symmetry applied
 * (random_symmetry, InflexionGame.py:115-122): sym[i] = k + 6*shift + 42*axis, k = rotation 0..5,
 * shift 0..6, axis 0..2 = 'r','q','s'; NULL = identity.  dtype: 0 = float32, 1 = int64. out is [n][4][7][7]. */
IFX_API int ifx_rules_planes(const ifx_state* states, const int32_t* sym, int64_t n, int dtype, void* out,
                             void* stream);

/* symmetries (InflexionGame.py:102-113): the 36 isomorphic forms of each [channels][7][7] array of
 * elem_size-byte elements (4 or 8), in the reference's order.  out is [n][36][channels][7][7]. */
IFX_API int ifx_rules_symmetries(const void* in, int64_t n, int channels, int elem_size, void* out,
                                 void* stream);

/* rotate / translate / random_symmetry of arbitrary board-like arrays (InflexionGame.py:115-196):
 * out[i] = translate(rotate(in[i], k), shift, axis) with sym[i] = k + 6*shift + 42*axis (0 <= sym < 126). */
IFX_API int ifx_rules_transform(const void* in, const int32_t* sym, int64_t n, int channels, int elem_size,
                                void* out, void* stream);

/* score / piece_count_diff / power_diff / total_power (InflexionGame.py:198-199, 312-323, 347):
 * info[i] = { score for side to move, power_diff for side to move, total_power, number of valid actions }. */
IFX_API int ifx_rules_info(const ifx_state* states, int64_t n, int32_t* info /* [n][4] */, void* stream);

/* RandomPlayer.play for every state (InflexionPlayers.py:24-34): a uniform draw over the valid actions in
 * ascending order, from Philox(seed; states[i].tag, step, 0, 3); -1 for finished games / no valid action. */
IFX_API int ifx_rules_random_actions(const ifx_state* states, int64_t n, uint64_t seed, uint32_t step,
                                     int32_t* actions, void* stream);

/* GreedyPlayer.play for every state (InflexionPlayers.py:61-77): best one-ply score(), ties to the highest
 * action index; -1 for finished games. */
IFX_API int ifx_rules_greedy_actions(const ifx_state* states, int64_t n, int32_t* actions, void* stream);

/* Host-buffer forms of the above (plain host pointers; copies in, launches, copies out, synchronises).
 * These are what the object-per-game drop-in API (InflexionGame.to_next_state etc.) calls. */
IFX_API int ifx_rules_mask_step_host(const ifx_state* states_in, const int32_t* actions, int64_t n,
                                     uint64_t* masks, ifx_state* states_out, int32_t* status);
IFX_API int ifx_rules_planes_host(const ifx_state* states, const int32_t* sym, int64_t n, int dtype, void* out);
IFX_API int ifx_rules_symmetries_host(const void* in, int64_t n, int channels, int elem_size, void* out);
IFX_API int ifx_rules_transform_host(const void* in, const int32_t* sym, int64_t n, int channels, int elem_size,
                                     void* out);
IFX_API int ifx_rules_info_host(const ifx_state* states, int64_t n, int32_t* info);

/* ---- batched MCTS: one warp per game tree ----------------------------------------------------- */

typedef struct ifx_mcts_config {
  int32_t n_games;          /* concurrent game trees on this GPU                                  */
  int32_t nodes_per_game;   /* node pool capacity per game (<= 16383)                             */
  int32_t hash_per_game;    /* hash slots per game, power of two, >= 2*nodes_per_game             */
  int32_t max_depth;        /* longest search path accepted, 1..352 (a game has at most 344 plies, so 352 = no
                               limit, as in the reference's recursion; 32 levels live in registers, deeper ones in
                               the workspace).  A deeper descent sets IFX_F_DEPTH                                   */
  int32_t num_sims;         /* args.numMCTSSims (MCTS.py:45), < 2^17                              */
  float cpuct;              /* float32(args.cpuct) (MCTS.py:122; NEP-50 weak scalar)              */
  uint64_t seed;            /* Philox key for every random draw on the path    */
  int32_t policy_is_logits; /* 0: nn_pi holds float32 probabilities (NNet.py:94 exp(log_softmax)), nn_v values;
                               2: nn_pi points at the net's last GEMM output, binary16 rows of pi_stride
                                  (>= 344, multiple of 4) halfs = 343 policy logits + the value pre-activation;
                                  softmax and tanh are fused into the expansion, nn_v is ignored           */
  int32_t terminal_sims_per_wave; /* 0: no limit.  T >= 1: after T simulations that ended in a terminal state
                               (no net needed, MCTS.py:85-87) the game yields until the next wave.  A game whose best
                               child is a won position would otherwise run ALL its remaining simulations inside one
                               launch while the other warps have long finished (the tail of the wave).  Results do
                               not depend on T; num_sims+1 waves still complete a move (see ifx_mcts_wave)            */
  float dirichlet_alpha;    /* optional root noise, NOT part of the reference (its MCTS has none): when            */
  float dirichlet_eps;      /* dirichlet_eps > 0, P(root,.) <- (1-eps) P + eps Dir(alpha) over the valid actions,
                               once per move, drawn on the device from Philox.  0 (default) = reference behaviour.
                               A root expanded by a wave yields until the next wave (its priors are perturbed in
                               between), so a move then needs num_sims+2 calls of ifx_mcts_wave                     */
  int32_t leaves_per_wave;  /* 0 or 1: the reference's sequential search (one leaf per game and wave; the parity mode).
                               K = 2..16: leaf-parallel PERFORMANCE mode with virtual loss (csrc/ifx_tree_vl.cuh): a game
                               collects up to K leaves per wave, row g*K + j of nn_pi / nn_v / nn_in belongs to leaf j of
                               game g (the buffers have n_games*K rows), a move needs about num_sims/K net calls; keep
                               calling ifx_mcts_wave until ifx_mcts_status shows sims_done == num_sims for every game.
                               Visit counts differ from the sequential search by construction.  (-1: the leaf-parallel
                               kernel with K = 1, which must reproduce the sequential search bit for bit -- a test hook) */
} ifx_mcts_config;

typedef struct ifx_mcts ifx_mcts;

/* Offsets (bytes) of the SoA pools inside the workspace, for inspection by tests (MCTS tables
 * Qsa/Nsa/Ns/Ps of MCTS.py:25-31 are not read by any caller of the reference). */
typedef struct ifx_mcts_layout {
  int64_t rows;       /* float [G][C][3][352]: P, N word (count | child hint | kind; see ifx_tree.cuh), Q.
                         COMPACT: entry i of a row belongs to the node's i-th valid action (ascending)    */
  int64_t q64;        /* double [G][C][344]: Q when its kind is Python-float; compact */
  int64_t node_key;   /* uint64 [G][C][2]                                                               */
  int64_t node_head;  /* {int32 Ns, float sqrt(Ns), float sqrt(Ns+1e-8), pad} [G][C]                    */
  int64_t hash;       /* uint64 [G][H][2]                                                               */
  int64_t free_stack; /* uint16 [G][C]                                                                  */
  int64_t per_game;   /* struct per game, see csrc/ifx_tree.cuh                                         */
  int64_t path;       /* uint32x4 [G][max_depth rounded up to 32]: {slot<<9|row entry, N word, Q bits, Ns}/level */
  int64_t roots;      /* ifx_state [G]                                                                  */
  int64_t leaves;     /* ifx_state [G]                                                                  */
  int64_t sym;        /* uint8 [G][max(num_sims,128) rounded up to 32][64]: per simulation of the current move, the source
                         cell of each destination cell under its random_symmetry (byte 63: the symmetry code)  */
  int64_t total;      /* == ifx_mcts_workspace_bytes                                                    */
} ifx_mcts_layout;

IFX_API int ifx_mcts_get_layout(const ifx_mcts_config* cfg, ifx_mcts_layout* out);
IFX_API size_t ifx_mcts_workspace_bytes(const ifx_mcts_config* cfg);

/* MCTS.__init__ (MCTS.py:21-31).  `workspace` is a device buffer of ifx_mcts_workspace_bytes, 256-B aligned. */
IFX_API int ifx_mcts_create(const ifx_mcts_config* cfg, void* workspace, size_t workspace_bytes, ifx_mcts** out);
IFX_API void ifx_mcts_destroy(ifx_mcts* m);
/* args.numMCTSSims (< 2^17) and float32(args.cpuct) for the waves launched from now on.  The values travel as kernel
 * arguments: a CUDA graph captured before the call keeps the old ones and must be captured again. */
IFX_API int ifx_mcts_set_params(ifx_mcts* m, int32_t num_sims, float cpuct);
/* CUDA-event timing of the wave kernel on its launching stream: while enabled, every ifx_mcts_wave (outside
 * graph capture) is bracketed by an event pair recorded inside the call; ifx_mcts_wave_time synchronises on them,
 * returns the summed kernel time and the number of launches (at most 4096 per read-out), and clears the ring. */
IFX_API int ifx_mcts_wave_timing(ifx_mcts* m, int enable);
IFX_API int ifx_mcts_wave_time(ifx_mcts* m, double* total_ms, int32_t* launches);
/* Optional per-game timing trace of ifx_mcts_wave: buf = device uint64 [G][8] receiving %globaltimer stamps
 * {start, first loads done, leaf expanded, value backed up, first tree level done, end}, the deepest level
 * reached and the number of hash probes; NULL switches it off (the default). */
IFX_API int ifx_mcts_set_debug(ifx_mcts* m, uint64_t* buf);

/* Empty trees + install roots (MCTS.reset, MCTS.py:147-148, followed by the first getActionProb).
 * roots[g].tag is the game id used as the RNG stream coordinate.  `which` (nullable, int32 [G]) selects
 * the games to reset; others are untouched. */
IFX_API int ifx_mcts_reset_games(ifx_mcts* m, const ifx_state* roots, const int32_t* which, void* stream);

/* Start the next getActionProb on new roots while keeping the trees (the reference never clears its
 * tables between moves: Coach.py:110, InflexionPlayers.py:86-88).  Nodes whose turn plane makes them
 * unreachable from the new root are reclaimed. */
IFX_API int ifx_mcts_set_roots(ifx_mcts* m, const ifx_state* roots, void* stream);

/* One lock-step wave over all games (MCTS.search, MCTS.py:62-145):
 *   1. if the game has a pending leaf, expand it from nn_pi[g] / nn_v[g] (MCTS.py:89-112) and back the
 *      value up its path (MCTS.py:134-145);
 *   2. run further simulations -- PUCT descent (MCTS.py:114-132), terminal leaves backed up at once
 *      (MCTS.py:85-87) -- until a simulation reaches an unexpanded state or num_sims are done;
 *   3. the unexpanded state's planes, under the random symmetry of MCTS.py:91-92, are written to
 *      nn_in[g] as float32 [4][7][7] and the game waits for the net.
 * A move needs at most num_sims net evaluations, i.e. num_sims+1 calls of this function (also with
 * terminal_sims_per_wave = T >= 1: after wave w, simulations done + leaf pending >= w+1, because a game yields only
 * after a leaf went to the net or after at least one terminal simulation of that wave). */
IFX_API int ifx_mcts_wave(ifx_mcts* m, const float* nn_pi, int64_t pi_stride, const float* nn_v, float* nn_in,
                          void* stream);

/* The same wave for the games [first_game, first_game + n_games) only; row g of nn_pi / nn_v / nn_in still belongs to
 * game g.  Games are independent, so disjoint parts may run on different streams: with two halves, the tree wave of one
 * half fills the issue slots the other half's net forward leaves idle (SelfPlayEngine(streams=2)). */
IFX_API int ifx_mcts_wave_part(ifx_mcts* m, int32_t first_game, int32_t n_games, const float* nn_pi, int64_t pi_stride,
                               const float* nn_v, float* nn_in, void* stream);

/* Root tables after the search (getActionProb, MCTS.py:48-49): any output may be NULL.
 * counts int32 [G][343]; q32 float [G][343]; q64 double [G][343]; qkind uint8 [G][343] (1 = float32 kind,
 * 0 = Python-float kind); ns int32 [G]. */
IFX_API int ifx_mcts_root_stats(ifx_mcts* m, int32_t* counts, float* q32, double* q64, uint8_t* qkind,
                                int32_t* ns, void* stream);

/* Action choice of Coach.executeEpisode (Coach.py:66-81 with MCTS.py:51-60): temp = (move+1 < temp_threshold);
 * temp 0 = uniform among arg-max counts, temp 1 = inverse-CDF sample of counts/sum in float64.
 * actions int32 [G] (-1 for games that are over). */
IFX_API int ifx_mcts_sample_actions(ifx_mcts* m, int32_t temp_threshold, int32_t* actions, void* stream);

/* game = game.to_next_state(action) for every game (Coach.py:82) and start the next move. */
IFX_API int ifx_mcts_advance(ifx_mcts* m, const int32_t* actions, void* stream);

/* Continuous self-play (Coach.learn starts the next episode when one ends, Coach.py:108-112): every game whose
 * root is terminal gets the start position, an empty tree and the RNG stream tag + tag_stride; *count (nullable,
 * device int32) is incremented once per restarted game. */
IFX_API int ifx_mcts_restart_finished(ifx_mcts* m, uint32_t tag_stride, int32_t* count, void* stream);

/* status int32 [G][8]: sims_done, pending, fault bits, move index, root outcome, live nodes, free nodes, 0. */
IFX_API int ifx_mcts_status(ifx_mcts* m, int32_t* status, void* stream);
/* roots out: ifx_state [G] (current root of every game). */
IFX_API int ifx_mcts_get_roots(ifx_mcts* m, ifx_state* roots, void* stream);
/* counters uint64 [4]: node visits V (expanded nodes traversed), expansions E, terminal leaves, simulations. */
IFX_API int ifx_mcts_counters(ifx_mcts* m, uint64_t* counters, void* stream);

/* ---- leaf gather: one contiguous batch for the policy/value net ------------------------------- */

/* The games that will search in the coming move -- root not terminal, simulations left, no fault -- in ascending
 * order: live_idx int32 [G] (first *count entries valid), count int32 [1] (device).  The reference calls the net once
 * per real leaf only (MCTS.py:89-93); with this list and the two row movers below the batched net forward runs on the
 * live games' rows instead of on all G (episode tails, Arena).  Call after reset / set_roots / advance. */
IFX_API int ifx_mcts_live_rows(ifx_mcts* m, int32_t* live_idx, int32_t* count, void* stream);

/* dst[k] = src[idx[k]] for k < *n_ptr; rows *n_ptr <= k < rows_cap of dst are zeroed.  Rows are row_bytes long
 * (a multiple of 4; the copy runs 16 bytes at a time when rows, strides and bases are multiples of 16), strides in
 * bytes, all pointers device pointers; n_ptr is read
 * on the DEVICE so a captured CUDA graph serves any live count up to rows_cap.
 * Use: nn_in rows [G][4][7][7] float32 of the live games -> packed [rows_cap][4][7][7]. */
IFX_API int ifx_gather_pack_rows(const void* src, int64_t src_stride, const int32_t* idx, const int32_t* n_ptr,
                                 int64_t rows_cap, int64_t row_bytes, void* dst, int64_t dst_stride, void* stream);
/* dst[idx[k]] = src[k] for k < *n_ptr (the inverse: the net's packed output rows back to the games' rows). */
IFX_API int ifx_gather_scatter_rows(const void* src, int64_t src_stride, const int32_t* idx, const int32_t* n_ptr,
                                    int64_t rows_cap, int64_t row_bytes, void* dst, int64_t dst_stride, void* stream);

/* ---- deterministic evaluator ------------------------------------------------------------------ */

/* The deterministic policy/value stub of the parity contract ("visit counts bit-exact ... when both use a fixed
 * deterministic net stub", BASELINE.json north_star) as one kernel: stubnet.py:stub_policy_value bit for bit.
 * nn_in float32 [n][4][7][7] (the planes NNetWrapper.predict receives, NNet.py:78-94), coef int64 [196] (the stub's
 * mixing coefficients), dyadic = 1 for integer priors / 0 for k/1000003 priors; pi float32 rows of pi_stride >= 343,
 * v float32 [n].  Used by the bit-exact tests and by tree-and-rules-only throughput runs (bench.py --net stub). */
IFX_API int ifx_stub_policy_value(const float* nn_in, int64_t n, const int64_t* coef, int dyadic, float* pi,
                                  int64_t pi_stride, float* v, void* stream);

/* ---- input stage of the policy/value net ------------------------------------------------------- */

/* im2col of the net input for conv1 (3x3, pad 1; inflexion/pytorch/InflexionNNet.py:24,41 of the reference):
 * planes float32 [n][4][7][7] -> out binary16 [n*49][48] (16-byte aligned), row = position*49 + y*7 + x, column =
 * c*9 + ky*3 + kx for the 36 taps (zero outside the board), columns 36..47 zero.  conv1 + BN + ReLU then is ONE
 * tensor-core GEMM [n*49,48] x [48,512] with a bias+ReLU epilogue, and its output is the channels-last activation
 * conv2 reads (NNetWrapper.predict_batch / predict_head, the batched form of NNet.py:78-94).  GPU build only. */
IFX_API int ifx_net_im2col3x3(const float* planes, int64_t n, void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IFX_B200_H */

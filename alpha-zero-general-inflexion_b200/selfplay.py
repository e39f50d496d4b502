"""Device-resident self-play for G concurrent Inflexion games on one GPU.

``SelfPlayEngine`` is the batched form of ``Coach.executeEpisode`` (Coach.py:41-90): per move it runs the
``numMCTSSims`` simulations of every game (tree kernel waves interleaved with batched net forwards), records
the training example of the move (root state + visit counts + side to move), samples the action
(temperature rule of Coach.py:66-68,81) and advances every game -- all on the device, no host round trip
inside a move.  Examples are expanded to their 36 symmetric forms and given their value targets at the end
(``examples()``), reproducing Coach.py:74-90 including the reference's player-alignment quirk
unless ``fix_alignment=True``.

Games are independent, so multi-GPU self-play is one engine per rank on its own shard of game ids with no
communication.
"""
from __future__ import annotations

import numpy as np
import torch

from . import rules
from .backend import Backend, default_backend
from .mcts_batched import BatchedMCTS
from .state import STATE_DTYPE, pack_states


class SelfPlayEngine:
    def __init__(self, n_games: int, num_sims: int, evaluate, cpuct=1.0, temp_threshold: int = 30, seed: int = 0,
                 max_turns: int = 343, game_id_base: int = 0, nodes_per_game: int | None = None,
                 record: bool = True, use_graph: bool | None = None, backend: Backend | None = None,
                 fused_head: bool = False, auto_restart: bool = False, restart_stride: int | None = None,
                 record_moves: int | None = None, dirichlet_alpha: float = 0.0, dirichlet_eps: float = 0.0,
                 streams: int = 1, leaves_per_wave: int = 1):
        """evaluate(nn_in f32 [G,4,7,7]) -> (pi f32 [G,343], v f32 [G]) on the device (e.g. nnet.predict_batch); with
        ``fused_head`` it is evaluate(nn_in, out f16 [G,344]) writing the raw head output (e.g. nnet.predict_head) and
        the tree kernel does the softmax / tanh.  ``auto_restart``: a game that ends starts its next episode at once
        (continuous self-play, Coach.py:108-112) with RNG stream id + ``restart_stride`` (default: n_games; pass the
        global game count when several ranks play).  ``streams`` = 2: inside the captured move the games are searched as
        two halves on two streams, so the tree wave of one half runs while the net evaluates the other half (the
        latency-bound wave fills issue slots the tensor-bound convolutions leave idle); results are identical -- games
        are independent.  Used with CUDA graphs and a dense batch only.  ``leaves_per_wave`` = K > 1: the leaf-parallel
        performance mode (virtual loss; BatchedMCTS) -- for small sets of games, where the sequential search leaves the
        net with tiny batches; one captured (net, wave) step replayed until the device reports the move done; not the
        reference's visit counts."""
        self.be = backend or default_backend()
        self.G, self.num_sims, self.temp_threshold = int(n_games), int(num_sims), int(temp_threshold)
        self.max_turns, self.game_id_base = int(max_turns), int(game_id_base)
        self.evaluate = evaluate
        self.mcts = BatchedMCTS(self.G, num_sims, cpuct, seed=seed, nodes_per_game=nodes_per_game, backend=self.be,
                                fused_head=fused_head, dirichlet_alpha=dirichlet_alpha, dirichlet_eps=dirichlet_eps,
                                leaves_per_wave=leaves_per_wave)
        self.leaf_parallel = int(leaves_per_wave) > 1
        self.record = record
        # a game lasts at most max_turns+1 moves; continuous play (auto_restart) records up to ``record_moves`` moves
        self.max_moves = int(record_moves or self.max_turns + 1)
        self.auto_restart = bool(auto_restart)
        self.restart_stride = int(restart_stride or self.G)
        self.restarts = self.be.zeros((1,), torch.int32)         # episodes finished and restarted so far
        if record:
            self.ex_states = self.be.empty((self.max_moves, self.G, 64), torch.uint8)
            self.ex_after = self.be.empty((self.max_moves, self.G, 64), torch.uint8)   # root after the move (its outcome)
            self.ex_counts = self.be.empty((self.max_moves, self.G, 343), torch.int32)
            self.ex_actions = self.be.empty((self.max_moves, self.G), torch.int32)
        want_graph = self.be.is_cuda if use_graph is None else bool(use_graph)
        self.use_graph = want_graph and not self.leaf_parallel
        # leaf-parallel mode: the number of waves of a move is decided on the device, so the captured unit is one
        # (net, wave) step (BatchedMCTS.search_captured) instead of the whole move
        self.capture_step = want_graph and self.leaf_parallel and self.be.is_cuda
        self._graphs = {}                                         # net batch size -> captured search of one move
        # fixed sets of episodes (no auto_restart) thin out as games end: the net then runs on the live games only
        # (BatchedMCTS.plan_rows: one host read per move, a captured graph per batch size)
        self.pack = not self.auto_restart and not self.leaf_parallel
        self.streams = 2 if (int(streams) == 2 and self.G >= 2) else 1
        self._side = None
        self.move = 0
        self.my_launches = 0                                      # kernels of libifx_b200 launched (bench evidence)
        self.extra_launches_per_search = 0                        # library kernels inside `evaluate` (the stub evaluator)
        self.restart()

    # ---- episode control ----------------------------------------------------------------------------
    def restart(self) -> None:
        """game.restarted() + a fresh MCTS for every game (Coach.py:109-111)."""
        roots = pack_states(np.zeros((self.G, 49), dtype=np.int64), 1, 0, 0, self.max_turns,
                            tags=self.game_id_base + np.arange(self.G))
        self.mcts.reset(roots)
        self.move = 0
        self.my_launches += 1

    def _wave_and_net(self) -> None:
        self.mcts.wave()
        self.mcts.feed(self.evaluate)

    def _capture(self) -> None:
        """The whole search of one move -- num_sims x (tree wave, net forward) + the closing wave -- as ONE CUDA
        graph: a single launch per move, no host involvement and no inter-launch gaps inside it."""
        s = torch.cuda.Stream(device=self.be.device)
        s.wait_stream(torch.cuda.current_stream(self.be.device))
        with torch.cuda.stream(s):
            rows = self.mcts.rows_active
            two = self.streams == 2 and rows == self.G
            half = self.G // 2
            # warm the allocator / cuDNN heuristics for every batch shape the capture will see (results unused)
            shapes = ((0, half), (half, self.G)) if two else ((0, rows),)
            for _ in range(2):
                for a, b in shapes:
                    if self.mcts.fused_head:
                        self.evaluate(self.mcts.nn_in[a:b], torch.empty_like(self.mcts.nn_head[a:b]))
                    else:
                        self.evaluate(self.mcts.nn_in[a:b])
        torch.cuda.current_stream(self.be.device).wait_stream(s)
        torch.cuda.synchronize(self.be.device)
        graph = torch.cuda.CUDAGraph()
        two = self.streams == 2 and self.mcts.rows_active == self.G
        if two and self._side is None:
            self._side = torch.cuda.Stream(device=self.be.device)
        with torch.cuda.graph(graph):
            if not two:
                for _ in range(self.mcts.waves_per_search - 1):
                    self._wave_and_net()
                self.mcts.wave()
            else:                                                # two independent chains: (wave, net)* for each half
                main = torch.cuda.current_stream(self.be.device)
                half = self.G // 2
                parts = ((0, half), (half, self.G - half))
                self._side.wait_stream(main)
                for stream, part in ((self._side, parts[1]), (main, parts[0])):
                    with torch.cuda.stream(stream):
                        for _ in range(self.mcts.waves_per_search - 1):
                            self.mcts.wave(part)
                            self.mcts.feed(self.evaluate, part)
                        self.mcts.wave(part)
                main.wait_stream(self._side)
        self._graphs[self.mcts.rows_active] = graph

    def search(self, wave_timer=None) -> None:
        """The numMCTSSims loop for all games.  ``wave_timer`` (optional): True launches every tree wave outside the
        CUDA graph (so the library can bracket it with events, see BatchedMCTS.wave_timing); a callable wraps each
        launch instead."""
        if self.leaf_parallel:                                   # waves until the device reports every game done
            self.mcts.search(self.evaluate, capture=self.capture_step)
            self.my_launches += self.mcts.waves_per_search
            return
        if wave_timer is True:
            wave_timer = lambda fn: fn()  # noqa: E731
        if self.pack and self.mcts.plan_rows() and self.mcts.live_count == 0:
            return                                               # every game is over: nothing to search
        if self.use_graph and wave_timer is None:
            if getattr(self.mcts, "params_version", 0) != getattr(self, "_graphs_version", 0):
                self._graphs.clear()                             # set_params since the capture: the graphs hold the old values
                self._graphs_version = getattr(self.mcts, "params_version", 0)
            if self.mcts.rows_active not in self._graphs:
                self._capture()
            self._graphs[self.mcts.rows_active].replay()
        elif wave_timer is None:
            for _ in range(self.mcts.waves_per_search - 1):
                self._wave_and_net()
            self.mcts.wave()
        else:
            for _ in range(self.mcts.waves_per_search - 1):
                wave_timer(self.mcts.wave)
                self._net_only()
            wave_timer(self.mcts.wave)
        self.my_launches += self.mcts.waves_per_search + self.extra_launches_per_search

    def _net_only(self) -> None:
        """Net forward as plain stream launches (timed mode): the stream stays continuously fed, so the event pair
        around the following tree wave brackets the kernel and not the turnaround after a graph launch."""
        self.mcts.feed(self.evaluate)

    def finish_move(self) -> torch.Tensor:
        """Record the example of this move, choose the action, play it.  Returns the actions (int32 [G], -1 = over)."""
        m = self.mcts
        if self.record and self.move < self.max_moves:
            self.be.check(self.be.dll.ifx_mcts_get_roots(m._h, self.ex_states[self.move].data_ptr(), self.be.stream()))
            self.be.check(self.be.dll.ifx_mcts_root_stats(m._h, self.ex_counts[self.move].data_ptr(), None, None, None,
                                                          None, self.be.stream()))
            acts = self.ex_actions[self.move]
            self.be.check(self.be.dll.ifx_mcts_sample_actions(m._h, self.temp_threshold, acts.data_ptr(), self.be.stream()))
            self.my_launches += 3
        else:
            acts = m.sample_actions(self.temp_threshold)
            self.my_launches += 1
        m.advance(acts)
        self.my_launches += 1
        if self.record and self.move < self.max_moves:
            self.be.check(self.be.dll.ifx_mcts_get_roots(m._h, self.ex_after[self.move].data_ptr(), self.be.stream()))
            self.my_launches += 1
        if self.auto_restart:
            m.restart_finished(self.restart_stride, self.restarts)
            self.my_launches += 1
        self.move += 1
        return acts

    def play_move(self, wave_timer=None) -> torch.Tensor:
        self.search(wave_timer)
        return self.finish_move()

    def play_episode(self, max_moves: int | None = None, collective: bool = False) -> int:
        """Play until every game is over (or ``max_moves``); returns the number of moves played.

        ``collective`` (under torch.distributed): the ranks decide TOGETHER when to stop -- an all-reduce of two flags
        every 16 moves -- so every rank plays the same number of moves (``gather_records`` needs equal shapes) and none
        plays on after all games everywhere have ended; a fault on any rank raises on every rank."""
        limit = self.max_moves if max_moves is None else min(max_moves, self.max_moves)
        dist = None
        if collective:
            import torch.distributed as dist_mod
            if dist_mod.is_available() and dist_mod.is_initialized() and dist_mod.get_world_size() > 1:
                dist = dist_mod
        while self.move < limit:
            self.play_move()
            if self.move % 16 == 0 or self.move == limit:
                st = self.mcts.status()
                flags = torch.stack([(st[:, 2] != 0).any(), (st[:, 4] == 0).any()]).to(torch.int32)   # [fault, someone plays]
                if dist is not None:
                    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
                fault, playing = (bool(x) for x in flags.tolist())
                if fault:
                    self.mcts.raise_on_faults()
                    raise RuntimeError("MCTS fault on another rank")
                if not playing:
                    break
        return self.move

    # ---- host-buffer step (the end-to-end form measured by bench.py) ---------------------------------
    def step_host(self, roots_host: torch.Tensor, counts_host: torch.Tensor, actions_host: torch.Tensor,
                  next_roots_host: torch.Tensor, new_episode_host: torch.Tensor | None = None) -> None:
        """getActionProb + action choice + to_next_state for G games given as pinned HOST buffers: copies the root
        states in, searches, copies visit counts, actions and next states out, and synchronises.
        ``new_episode_host`` (int32 [G], optional) marks games whose root starts a new episode: their trees are
        emptied first, as Coach.learn builds a fresh MCTS per episode (Coach.py:109-110)."""
        dev_roots = roots_host.to(self.be.device, non_blocking=True)
        self.mcts.set_roots(dev_roots)
        if new_episode_host is not None:
            self.mcts.reset(dev_roots, which=new_episode_host.to(self.be.device, non_blocking=True))
            self.my_launches += 1
        self.search()
        stats = self.mcts.root_stats()
        acts = self.mcts.sample_actions(self.temp_threshold)
        self.mcts.advance(acts)
        counts_host.copy_(stats["counts"], non_blocking=True)
        actions_host.copy_(acts, non_blocking=True)
        next_roots_host.copy_(self.mcts.roots(), non_blocking=True)
        self.my_launches += 5
        if self.be.is_cuda:
            torch.cuda.current_stream(self.be.device).synchronize()
        self.move += 1

    # ---- training examples (Coach.py:74-90) ----------------------------------------------------------
    def records(self) -> dict:
        """The un-expanded self-play record of this rank: per move and game slot the root state, the state after the
        move, the root visit counts and the action (1.5 KB per move; the 36-fold expansion happens in ``examples``)."""
        assert self.record
        n = min(self.move, self.max_moves)
        return {"states": self.ex_states[:n], "after": self.ex_after[:n], "counts": self.ex_counts[:n],
                "actions": self.ex_actions[:n], "temp_threshold": self.temp_threshold}

    def gather_records(self) -> dict:
        """The final example gather (BASELINE.json north_star): all ranks' un-expanded records
        concatenated along the game axis with one all-gather per field (NCCL on GPUs) -- never the expanded set,
        which is 36x larger.  Every rank must have played the same number of moves and games."""
        import torch.distributed as dist
        rec = self.records()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return rec
        world = dist.get_world_size()
        out = {"temp_threshold": rec["temp_threshold"]}
        for k in ("states", "after", "counts", "actions"):
            t = rec[k].contiguous()
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            out[k] = torch.cat(parts, dim=1)
        return out

    def examples(self, games=None, fix_alignment: bool = False, expand_symmetries: bool = True):
        """-> (planes int64 [E,4,7,7], pi float64 [E,343], v float64 [E]) for every FINISHED episode of the selected
        game slots of this rank; see ``build_examples``."""
        return build_examples(self.be, self.records(), games, fix_alignment, expand_symmetries)


def build_examples(be, rec: dict, games=None, fix_alignment: bool = False, expand_symmetries: bool = True):
    """Training examples from a self-play record (``SelfPlayEngine.records`` / ``gather_records``) ->
    (planes int64 [E,4,7,7], pi float64 [E,343], v float64 [E]) for every FINISHED episode of the selected game
    slots, in the reference's order: per episode, per move, the 36 symmetric forms (Coach.py:77-78,89-90).

    Value targets follow Coach.py:89-90.  The reference extends ``players`` by the cumulative example count
    (Coach.py:79), so entry i of the zip pairs example i with players[i] where players = [P1]*36 + [P2]*72 +
    [P1]*108 + ...; that quirk is reproduced unless ``fix_alignment``."""
    ex_states, ex_after, ex_counts, ex_actions = rec["states"], rec["after"], rec["counts"], rec["actions"]
    temp_threshold = int(rec["temp_threshold"])
    n_moves, G = ex_actions.shape[0], ex_actions.shape[1]
    dev = be.device
    st_all = ex_states.cpu().numpy().view(STATE_DTYPE).reshape(n_moves, G)
    after_all = ex_after.cpu().numpy().view(STATE_DTYPE).reshape(n_moves, G)
    acts_all = ex_actions.cpu().numpy()
    games = list(range(G) if games is None else games)
    per = 36 if expand_symmetries else 1
    value_of = {0: 0.0, 1: 1e-4, 2: 1.0, 3: -1.0}
    # 1. finished episodes: (slot, first move, last move); value targets per example (host, tiny)
    sel_mv, sel_g, values = [], [], []
    for g in games:
        start = 0
        for mv in range(n_moves):
            if acts_all[mv, g] < 0:                       # slot idle (game over, no restart)
                start = mv + 1
                continue
            if after_all[mv, g]["outcome"] == 0:
                continue
            rows = np.arange(start, mv + 1)               # one finished episode: moves start..mv of slot g
            start = mv + 1
            played = len(rows)
            movers = np.where(st_all[rows, g]["player"] == 0, 1, -1)   # side to move of each recorded state
            final = after_all[mv, g]
            end_player = 1 if final["player"] == 0 else -1
            result = value_of[int(final["outcome"])]
            if fix_alignment or per == 1:
                players = np.repeat(movers, per)
            else:
                players = np.concatenate([np.full(per * (i + 1), movers[i]) for i in range(played)])[: per * played]
            values.append(np.where(players == end_player, result, -result).astype(np.float64))
            sel_mv.append(rows)
            sel_g.append(np.full(played, g))
    if not sel_mv:
        return np.zeros((0, 4, 7, 7), np.int64), np.zeros((0, 343)), np.zeros(0)
    mv_idx = torch.from_numpy(np.concatenate(sel_mv)).to(dev)
    g_idx = torch.from_numpy(np.concatenate(sel_g)).to(dev)
    # 2. planes, pi and their 36 symmetric forms for ALL selected (move, game) pairs in a handful of kernel calls
    states = ex_states[mv_idx, g_idx].contiguous()
    planes = rules.planes(states, dtype=torch.int64, backend=be)
    counts = ex_counts[mv_idx, g_idx].to(torch.float64)
    acts = ex_actions[mv_idx, g_idx].long()
    ep_move = torch.from_numpy(np.concatenate([np.arange(len(r)) for r in sel_mv])).to(dev)
    pi = counts / counts.sum(dim=1, keepdim=True)                                    # MCTS.py:58-60 at temp 1
    onehot = torch.zeros_like(pi)
    onehot[torch.arange(len(acts), device=dev), acts] = 1.0                          # MCTS.py:51-56 at temp 0
    pi = torch.where(((ep_move + 1) < temp_threshold)[:, None], pi, onehot)
    if expand_symmetries:
        planes = rules.symmetries(planes.contiguous(), backend=be).reshape(-1, 4, 7, 7)
        pi = rules.symmetries(pi.reshape(-1, 7, 7, 7).contiguous(), backend=be).reshape(-1, 343)
    return planes.cpu().numpy(), pi.cpu().numpy(), np.concatenate(values)


def _sym36_code(form: np.ndarray) -> np.ndarray:
    """Index 0..35 in the order of InflexionGame.symmetries (InflexionGame.py:102-113) -> symmetry code k + 6*shift
    (axis 'r'): identity, rotations 1..5, then for k = 1..5 the translations j = 1..6 of rotation k."""
    form = np.asarray(form)
    k, j = (form - 6) // 6 + 1, (form - 6) % 6 + 1
    return np.where(form < 6, form, k + 6 * j).astype(np.int32)


class ExampleSet:
    """The training examples of a self-play record WITHOUT materialising them.

    ``build_examples`` expands every recorded move into its 36 symmetric forms (1.6 KB each as arrays, far more as the
    reference's Python lists): 19 MB per game, 637 GB for BASELINE config 4's 32 768 games.  An
    ``ExampleSet`` keeps the un-expanded record (1.5 KB per move) plus an index of the finished episodes and produces
    any batch of expanded examples on demand -- ``batch(idx)`` returns exactly rows ``idx`` of what ``build_examples``
    would return (same order, same symmetry, same value targets incl. the reference's alignment quirk D1), computed
    for that batch only by the rules kernels.  ``NNetWrapper.train`` accepts it in place of the example list, which
    makes its uniform sampling with replacement (NNet.py:52) the same distribution at O(record) memory."""

    def __init__(self, be, rec: dict, games=None, fix_alignment: bool = False):
        self.be = be
        self.temp_threshold = int(rec["temp_threshold"])
        ex_states, ex_after, ex_counts, ex_actions = rec["states"], rec["after"], rec["counts"], rec["actions"]
        n_moves, G = ex_actions.shape[0], ex_actions.shape[1]
        st_all = ex_states.cpu().numpy().view(STATE_DTYPE).reshape(n_moves, G)
        after_all = ex_after.cpu().numpy().view(STATE_DTYPE).reshape(n_moves, G)
        acts_all = ex_actions.cpu().numpy()
        value_of = {0: 0.0, 1: 1e-4, 2: 1.0, 3: -1.0}
        mv, gg, ep_move, ep_first, ep_len, mover, end_player, result = [], [], [], [], [], [], [], []
        total = 0
        for g in (range(G) if games is None else games):
            start = 0
            for m in range(n_moves):
                if acts_all[m, g] < 0:
                    start = m + 1
                    continue
                if after_all[m, g]["outcome"] == 0:
                    continue
                rows = np.arange(start, m + 1)
                start = m + 1
                n = len(rows)
                final = after_all[m, g]
                mv.append(rows); gg.append(np.full(n, g)); ep_move.append(np.arange(n))
                ep_first.append(np.full(n, total)); ep_len.append(np.full(n, n))
                mover.append(np.where(st_all[rows, g]["player"] == 0, 1, -1))
                end_player.append(np.full(n, 1 if final["player"] == 0 else -1))
                result.append(np.full(n, value_of[int(final["outcome"])]))
                total += n
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
        self.mv, self.g, self.ep_move = cat(mv, np.int64), cat(gg, np.int64), cat(ep_move, np.int64)
        self.ep_first, self.ep_len = cat(ep_first, np.int64), cat(ep_len, np.int64)
        self.mover, self.end_player, self.result = cat(mover, np.int64), cat(end_player, np.int64), cat(result, np.float64)
        self.fix_alignment = bool(fix_alignment)
        # only the selected (move, game) pairs are kept, on the device: states 64 B + counts 1 372 B + action 4 B each
        dev = be.device
        mi, gi = torch.from_numpy(self.mv).to(dev), torch.from_numpy(self.g).to(dev)
        self.states = ex_states[mi, gi].contiguous()
        self.counts = ex_counts[mi, gi].contiguous()
        self.actions = ex_actions[mi, gi].contiguous()

    def __len__(self) -> int:
        return 36 * len(self.mv)

    def _values(self, rec_idx: np.ndarray, form: np.ndarray) -> np.ndarray:
        """Value target of expanded example (record rec_idx, symmetric form) -- Coach.py:89-90 with players[] as the
        reference builds it (Coach.py:79: extended by the CUMULATIVE example count)."""
        if self.fix_alignment:
            player = self.mover[rec_idx]
        else:
            li = 36 * self.ep_move[rec_idx] + form                     # position inside the episode's example list
            # players = [mover_0]*36 + [mover_1]*72 + [mover_2]*108 + ...: entry li belongs to the j with
            # 18 j (j+1) <= li < 18 (j+1)(j+2)
            j = ((np.sqrt(1.0 + 4.0 * li / 18.0) - 1.0) / 2.0).astype(np.int64)
            j = np.where(18 * (j + 1) * (j + 2) <= li, j + 1, j)
            j = np.where(18 * j * (j + 1) > li, j - 1, j)
            player = self.mover[self.ep_first[rec_idx] + j]
        return np.where(player == self.end_player[rec_idx], self.result[rec_idx], -self.result[rec_idx])

    def batch(self, idx, planes_dtype=torch.float32, target_dtype=torch.float32):
        """Rows ``idx`` (any order, repeats allowed) of the expanded example list ->
        (planes [B,4,7,7] planes_dtype, pi target_dtype [B,343], v target_dtype [B]) on the device."""
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        rec_idx, form = idx // 36, idx % 36
        dev = self.be.device
        ri = torch.from_numpy(rec_idx).to(dev)
        sym = torch.from_numpy(_sym36_code(form)).to(dev)
        planes = rules.planes(self.states[ri].contiguous(), sym=sym, dtype=planes_dtype, backend=self.be)
        counts = self.counts[ri].to(torch.float64)
        pi = counts / counts.sum(dim=1, keepdim=True)                                    # MCTS.py:58-60 at temp 1
        onehot = torch.zeros_like(pi)
        onehot[torch.arange(len(idx), device=dev), self.actions[ri].long()] = 1.0          # MCTS.py:51-56 at temp 0
        temp1 = torch.from_numpy((self.ep_move[rec_idx] + 1) < self.temp_threshold).to(dev)
        pi = torch.where(temp1[:, None], pi, onehot)
        pi = rules.transform(pi.reshape(-1, 7, 7, 7).contiguous(), sym, backend=self.be).reshape(-1, 343)
        v = torch.from_numpy(self._values(rec_idx, form)).to(dev)
        return planes, pi.to(target_dtype), v.to(target_dtype)

    def materialise(self):
        """The full expanded set as ``build_examples`` returns it (tests; small sets only)."""
        planes, pi, v = self.batch(np.arange(len(self)), planes_dtype=torch.int64, target_dtype=torch.float64)
        return planes.cpu().numpy(), pi.cpu().numpy(), v.cpu().numpy()


class ExampleUnion:
    """Several ``ExampleSet``s (the iterations of Coach.trainExamplesHistory) addressed as one example list."""

    def __init__(self, sets):
        self.sets = [s for s in sets if len(s)]
        self.offsets = np.cumsum([0] + [len(s) for s in self.sets])

    def __len__(self) -> int:
        return int(self.offsets[-1])

    def batch(self, idx, planes_dtype=torch.float32, target_dtype=torch.float32):
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        which = np.searchsorted(self.offsets, idx, side="right") - 1
        outs, order = [], []
        for k in np.unique(which):
            sel = np.nonzero(which == k)[0]
            outs.append(self.sets[k].batch(idx[sel] - self.offsets[k], planes_dtype, target_dtype))
            order.append(sel)
        inv = torch.from_numpy(np.argsort(np.concatenate(order))).to(outs[0][0].device)
        return tuple(torch.cat([o[i] for o in outs])[inv] for i in range(3))

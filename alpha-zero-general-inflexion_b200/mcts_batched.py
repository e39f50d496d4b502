"""Batched MCTS over G concurrent game trees on one GPU (one warp per tree).

Host-side driver of the ifx_mcts_* kernels (include/ifx_b200.h).  One instance replaces G instances of the
reference's ``MCTS`` (MCTS.py:15-148) together with the per-move part of ``Coach.executeEpisode``
(Coach.py:66-82): ``search`` = the ``numMCTSSims`` loop of ``getActionProb`` (MCTS.py:45-46) for all games in
lock step, ``root_stats`` = the visit counts of MCTS.py:48-49, ``sample_actions`` = MCTS.py:51-60 + Coach.py:81,
``advance`` = Coach.py:82.

The policy/value net is a callable ``evaluate(nn_in f32 [G,4,7,7]) -> (pi f32 [G,343], v f32 [G])`` working
on device tensors; per move it is called at most ``num_sims`` times with the full batch.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from .backend import Backend, default_backend


def default_pool_size(num_sims: int) -> int:
    """Nodes per game.  Live nodes (turn > root turn) were measured at <= 2.8 x sims; the bound
    is depth x sims, so keep a wide margin and fail loudly (IFX_F_POOL) rather than silently."""
    n = 64
    while n < 10 * max(num_sims, 1):
        n *= 2
    return min(n, 16383)


NODE_BYTES = 3 * 352 * 4 + 344 * 8 + 16 + 16 + 2 + 2 * 2 * 16      # rows + q64 + key + head + free-stack entry + hash share


def auto_pool_size(n_games: int, num_sims: int, device, fraction: float = 0.5) -> int:
    """Nodes per game for ``nodes_per_game="auto"``: as many as ``fraction`` of the device's FREE memory holds, between
    ``default_pool_size`` and the format's limit of 16383.  The reference never frees a node; here a node is reclaimed
    once the root's turn has passed its own (the only reclamation that cannot change a result: an unreachable-looking
    node with a later turn can still be found again through a transposition).  A confident net keeps its principal
    variation -- about ``num_sims`` new nodes per move -- alive for as many moves as the variation is deep, so the live
    set can reach tens of times ``num_sims``; small batches (Coach, Arena, one-game MCTS objects) can afford that."""
    floor = default_pool_size(num_sims)
    dev = torch.device(device)
    if dev.type != "cuda":
        return floor
    free, _ = torch.cuda.mem_get_info(dev)
    return int(max(floor, min(16383, fraction * free / (max(n_games, 1) * NODE_BYTES))))


class BatchedMCTS:
    def __init__(self, n_games: int, num_sims: int, cpuct=1.0, seed: int = 0, nodes_per_game: int | str | None = None,
                 max_depth: int = 352, backend: Backend | None = None, fused_head: bool = False,
                 pad_policy_rows: bool = True, terminal_sims_per_wave: int = 1, dirichlet_alpha: float = 0.0,
                 dirichlet_eps: float = 0.0, pack_quantum: int | None = None, leaves_per_wave: int = 1):
        """fused_head: the net hands over its last GEMM output (float16 [G,344]: 343 policy logits + the value
        pre-activation, buffer ``nn_head``) and the tree kernel applies softmax / tanh itself.
        terminal_sims_per_wave: how many simulations that end in a terminal state (no net needed) a game may run in
        one wave before it yields (0 = no limit); results do not depend on it, the length of a wave's tail does.
        leaves_per_wave: 1 = the reference's sequential search (parity mode).  K > 1 = leaf-parallel PERFORMANCE mode with
        virtual loss (csrc/ifx_tree_vl.cuh): every game hands up to K leaves to each net call (the net buffers have
        n_games * K rows) and a move needs about num_sims / K calls; visit counts differ from the reference's by
        construction.  ``search`` then runs until the device reports every game done.
        pack_quantum: granularity of the packed net batch (see ``plan_rows``); None = 256 rows from 2048 games on, 32 rows
        from 64 games on, otherwise no packing; 0 = never pack.
        dirichlet_alpha / dirichlet_eps: optional root noise P <- (1-eps) P + eps Dir(alpha), drawn on the device once
        per move.  The reference's MCTS has none, so the default (eps = 0) is off and the bit-exact parity holds."""
        self.be = backend or default_backend()
        self.n_games, self.num_sims = int(n_games), int(num_sims)
        if nodes_per_game == "auto":
            nodes_per_game = auto_pool_size(self.n_games, num_sims, self.be.device)
        nodes = int(nodes_per_game or default_pool_size(num_sims))
        hash_slots = 64
        while hash_slots < 2 * nodes:
            hash_slots *= 2
        # float32(cpuct): the reference multiplies a Python number into a float32 (NEP-50 weak scalar, MCTS.py:122)
        self.cfg = _capi.MctsConfig(n_games=self.n_games, nodes_per_game=nodes, hash_per_game=hash_slots,
                                    max_depth=max_depth, num_sims=self.num_sims, cpuct=float(np.float32(cpuct)),
                                    seed=int(seed) & (2 ** 64 - 1), policy_is_logits=2 if fused_head else 0,
                                    terminal_sims_per_wave=int(terminal_sims_per_wave),
                                    dirichlet_alpha=float(dirichlet_alpha), dirichlet_eps=float(dirichlet_eps),
                                    leaves_per_wave=int(leaves_per_wave))
        self.leaves_per_wave = max(1, int(leaves_per_wave))
        self._adaptive = int(leaves_per_wave) > 1 or int(leaves_per_wave) == -1
        dll = self.be.dll
        self.layout = _capi.MctsLayout()
        self.be.check(dll.ifx_mcts_get_layout(C.byref(self.cfg), C.byref(self.layout)))
        nbytes = dll.ifx_mcts_workspace_bytes(C.byref(self.cfg))
        # the caller owns the workspace: one uint8 tensor (+256 B so the base can be aligned)
        self._ws_raw = self.be.empty((nbytes + 256,), torch.uint8)
        off = (-self._ws_raw.data_ptr()) % 256
        self.workspace = self._ws_raw[off:off + nbytes]
        handle = C.c_void_p()
        self.be.check(dll.ifx_mcts_create(C.byref(self.cfg), self.workspace.data_ptr(), nbytes, C.byref(handle)))
        self._h = handle
        g = self.n_games * self.leaves_per_wave                 # rows of the net buffers (K leaves per game in the leaf-parallel mode)
        # net I/O buffers (persistent so the wave + net step can be captured in a CUDA graph)
        self.nn_in = self.be.zeros((g, 4, 7, 7), torch.float32)
        if pad_policy_rows:
            self._nn_pi_rows = self.be.zeros((g, 352), torch.float32)    # rows padded to 352 floats: 16-byte aligned
            self.nn_pi = self._nn_pi_rows[:, :343]
        else:                                                        # dense [G,343] rows (the kernel's scalar-load path)
            self.nn_pi = self.be.zeros((g, 343), torch.float32)
        self.nn_v = self.be.zeros((g,), torch.float32)
        self.fused_head = bool(fused_head)
        self.dirichlet_eps = float(dirichlet_eps)
        self.nn_head = self.be.zeros((g, 344), torch.float16) if fused_head else None
        self.waves_run = 0
        self.evals_run = 0
        self.rows_evaluated = 0                                 # net rows evaluated so far (evals x active rows)
        # leaf gather: the games that search in the current move, and the packed net buffers per batch size
        g = self.n_games
        if pack_quantum is None:
            pack_quantum = 256 if g >= 2048 else (32 if g >= 64 else 0)
        self.pack_quantum = int(pack_quantum) if self.leaves_per_wave == 1 else 0     # (packing: sequential mode only)
        self._live_idx = self.be.zeros((g,), torch.int32)
        self._live_n = self.be.zeros((1,), torch.int32)
        self.live_count = g
        self.rows_active = g                                    # rows the net runs on in the current move (g = dense)
        self._packed = {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self.be.dll.ifx_mcts_destroy(h)
            except Exception:
                pass

    # ---- tree lifecycle -------------------------------------------------------------------------
    def reset(self, roots, which: torch.Tensor | None = None) -> None:
        """Empty the trees (of the games selected by ``which``) and install ``roots``; roots[g].tag = game id."""
        r = self.be.states_tensor(roots)
        assert r.shape == (self.n_games, 64)
        w = None if which is None else which.to(self.be.device, torch.int32).contiguous()
        self.be.check(self.be.dll.ifx_mcts_reset_games(self._h, r.data_ptr(), None if w is None else w.data_ptr(),
                                                       self.be.stream()))
        self._keep = (r, w)

    def set_roots(self, roots) -> None:
        """Next getActionProb on new root states, keeping the tables (the reference never clears them)."""
        r = self.be.states_tensor(roots)
        assert r.shape == (self.n_games, 64)
        self.be.check(self.be.dll.ifx_mcts_set_roots(self._h, r.data_ptr(), self.be.stream()))
        self._keep = (r, None)

    @property
    def waves_per_search(self) -> int:
        """ifx_mcts_wave calls that complete a move: num_sims + 1 (+1 with root noise: a freshly expanded root yields)."""
        return self.num_sims + 1 + (1 if self.dirichlet_eps > 0 else 0)

    def set_params(self, num_sims: int, cpuct) -> None:
        """args.numMCTSSims / args.cpuct for the searches that follow.  A CUDA graph captured earlier has the old values
        baked into its kernel arguments: ``params_version`` changes so that owners of such graphs (SelfPlayEngine)
        re-capture."""
        cp = float(np.float32(cpuct))
        if int(num_sims) == self.num_sims and cp == self.cfg.cpuct:
            return                                             # unchanged: captured graphs stay valid
        self.params_version = getattr(self, "params_version", 0) + 1
        self.num_sims, self.cfg.cpuct = int(num_sims), cp
        self.be.check(self.be.dll.ifx_mcts_set_params(self._h, self.num_sims, cp))

    # ---- search ---------------------------------------------------------------------------------
    def wave(self, part: tuple[int, int] | None = None) -> None:
        """One lock-step wave over all games, or over the games ``part = (first, count)`` only (ifx_mcts_wave_part)."""
        first, count = part if part is not None else (0, self.n_games)
        if self.fused_head:
            self.be.check(self.be.dll.ifx_mcts_wave_part(self._h, first, count, self.nn_head.data_ptr(), self.nn_head.stride(0),
                                                         None, self.nn_in.data_ptr(), self.be.stream()))
        else:
            self.be.check(self.be.dll.ifx_mcts_wave_part(self._h, first, count, self.nn_pi.data_ptr(), self.nn_pi.stride(0),
                                                         self.nn_v.data_ptr(), self.nn_in.data_ptr(), self.be.stream()))
        self.waves_run += 1

    # ---- leaf gather (north_star kernel 3) -----------------------------------------------------------
    def plan_rows(self) -> int:
        """Decide the net batch of the coming move.  Lists the games that will search (root not over, simulations left:
        ``ifx_mcts_live_rows``), reads their number (ONE host sync per move) and rounds it up to ``pack_quantum``.  When
        that is below the game count, every ``feed`` of the move packs the live games' planes into the first rows of a
        batch of that size, runs the net on it and scatters the answers back -- the reference evaluates one leaf per
        real simulation only (MCTS.py:89-93).  Returns the batch size (``n_games`` = dense, no packing).
        Call after reset / set_roots / advance; without a call the batch stays dense."""
        self.be.check(self.be.dll.ifx_mcts_live_rows(self._h, self._live_idx.data_ptr(), self._live_n.data_ptr(), self.be.stream()))
        self.live_count = int(self._live_n.item())
        q, g = self.pack_quantum, self.n_games
        rows = g if q <= 0 else min(g, max(q, -(-self.live_count // q) * q))
        self.rows_active = rows
        return rows

    def dense_rows(self) -> None:
        self.rows_active, self.live_count = self.n_games, self.n_games

    def _packed_buffers(self, rows: int):
        if rows not in self._packed:
            b = {"in": self.be.zeros((rows, 4, 7, 7), torch.float32)}
            if self.fused_head:
                b["head"] = self.be.zeros((rows, 344), torch.float16)
            else:
                b["pi"] = self.be.zeros((rows, 352), torch.float32)
                b["v"] = self.be.zeros((rows,), torch.float32)
            self._packed[rows] = b
        return self._packed[rows]

    def _move_rows(self, fn, src, dst, rows: int, row_bytes: int) -> None:
        self.be.check(fn(src.data_ptr(), src.stride(0) * src.element_size(), self._live_idx.data_ptr(), self._live_n.data_ptr(),
                         rows, row_bytes, dst.data_ptr(), dst.stride(0) * dst.element_size(), self.be.stream()))

    def _feed_packed(self, evaluate, rows: int) -> None:
        dll, b = self.be.dll, self._packed_buffers(rows)
        self._move_rows(dll.ifx_gather_pack_rows, self.nn_in, b["in"], rows, 196 * 4)
        if self.fused_head:
            evaluate(b["in"], b["head"])
            self._move_rows(dll.ifx_gather_scatter_rows, b["head"], self.nn_head, rows, 344 * 2)
        else:
            pi, v = evaluate(b["in"])
            b["pi"][:, :343].copy_(pi)
            b["v"].copy_(v.reshape(-1))
            padded = hasattr(self, "_nn_pi_rows")
            self._move_rows(dll.ifx_gather_scatter_rows, b["pi"], self._nn_pi_rows if padded else self.nn_pi, rows,
                            (344 if padded else 343) * 4)
            self._move_rows(dll.ifx_gather_scatter_rows, b["v"], self.nn_v, rows, 4)

    def feed(self, evaluate, part: tuple[int, int] | None = None) -> None:
        """One net evaluation of the pending leaves into the buffers the next wave reads (``part``: of the games
        (first, count) only -- dense mode)."""
        if part is not None:
            a, b = part[0], part[0] + part[1]
            self.rows_evaluated += part[1]
            if self.fused_head:
                evaluate(self.nn_in[a:b], self.nn_head[a:b])
            else:
                pi, v = evaluate(self.nn_in[a:b])
                if pi.data_ptr() != self.nn_pi[a:b].data_ptr():
                    self.nn_pi[a:b].copy_(pi)
                if v.data_ptr() != self.nn_v[a:b].data_ptr():
                    self.nn_v[a:b].copy_(v.reshape(-1))
            self.evals_run += 1
            return
        self.rows_evaluated += self.rows_active
        if self.rows_active < self.n_games:
            self._feed_packed(evaluate, self.rows_active)
            self.evals_run += 1
            return
        if self.fused_head:
            evaluate(self.nn_in, self.nn_head)              # writes float16 [G,344] in place
        else:
            pi, v = evaluate(self.nn_in)
            if pi is not self.nn_pi:
                self.nn_pi.copy_(pi)
            if v is not self.nn_v:
                self.nn_v.copy_(v.reshape(-1))
        self.evals_run += 1

    def search(self, evaluate, pack: bool = False, capture: bool = False, version=0) -> None:
        """All ``num_sims`` simulations of every game: num_sims x (wave, net) + a closing wave.  No host sync, unless
        ``pack``: then ``plan_rows`` first (one sync) so the net only sees the games that search in this move.
        ``capture`` (CUDA only): replay the (net, wave) step from a CUDA graph -- see ``search_captured``."""
        if capture and self.be.is_cuda:
            if pack and not self._adaptive:
                self.plan_rows()
            return self.search_captured(evaluate, version)
        if self._adaptive:
            return self.search_adaptive(evaluate)
        if pack:
            self.plan_rows()
        for _ in range(self.waves_per_search - 1):
            self.wave()
            self.feed(evaluate)
        self.wave()

    def search_adaptive(self, evaluate) -> None:
        """Same result as ``search`` but asks the device whether any game still waits for the net (one small read-back)
        and stops as soon as none does.  Sequential mode (the object-per-game drop-in, where a ``predict`` call is
        expensive): a read-back after every wave, so the net is called exactly as often as the reference calls
        ``predict``.  Leaf-parallel mode: a wave completes at most ``leaves_per_wave`` net-evaluated simulations per
        game, so after a look the next ceil(remaining / K) - 1 waves cannot be the last ones and run without a
        read-back."""
        self.wave()
        st = self.status()
        while True:
            done = (st[:, 0] >= self.num_sims) | (st[:, 2] != 0)
            if bool(done.all()):
                return
            if bool(st[:, 1].any()):                         # (a game that yielded on its terminal budget needs no net)
                self.feed(evaluate)
            if self.leaves_per_wave > 1:
                remaining = int((self.num_sims - st[:, 0])[~done].max())
                for _ in range(max(0, -(-remaining // self.leaves_per_wave) - 1)):
                    self.wave()
                    self.feed(evaluate)
            self.wave()
            st = self.status()

    # ---- captured step ---------------------------------------------------------------------------
    def _captured_step(self, evaluate, version):
        """CUDA graph of ONE search step -- net forward on the pending leaves, then the tree wave that consumes the
        answers -- captured on first use per net batch size, and again whenever ``version`` (the caller's token for the
        net's weights) or the search parameters change.  The kernel arguments baked into it are the persistent
        buffers of this object."""
        if not self.be.is_cuda:
            raise RuntimeError("search_captured needs the CUDA backend")
        key = (version, getattr(self, "params_version", 0))
        if getattr(self, "_steps_key", None) != key:
            self._steps, self._steps_key = {}, key             # (drops the old graphs and their memory pools)
        if self.rows_active in self._steps:
            return self._steps[self.rows_active]
        dev = self.be.device
        rows = self.rows_active if self.rows_active < self.n_games else self.nn_in.shape[0]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                          # warm the allocator / cuDNN heuristics for this batch shape
            x = self._packed_buffers(rows)["in"] if rows < self.n_games else self.nn_in
            for _ in range(2):
                if self.fused_head:
                    evaluate(x, torch.empty((rows, 344), dtype=torch.float16, device=dev))
                else:
                    evaluate(x)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        counters = (self.waves_run, self.evals_run, self.rows_evaluated)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.feed(evaluate)
            self.wave()
        self.waves_run, self.evals_run, self.rows_evaluated = counters      # (capturing launches nothing)
        self._steps[self.rows_active] = graph
        return graph

    def search_captured(self, evaluate, version=0) -> None:
        """``search`` with every (net, wave) step replayed from one CUDA graph: the same kernels on the same data in the
        same order, so the same result bit for bit, at one launch per simulation step instead of a dozen -- for the
        object-per-game drop-in and small sets of games, where the eager loop is bound by launch latency.  Sequential
        mode needs no read-back at all: ``num_sims`` steps complete every game (a wave on a finished game does
        nothing).  Leaf-parallel mode looks at the device after every ceil(remaining / K) steps.  ``evaluate`` must be
        capturable (device tensors in and out, no host sync) and must not change while ``version`` stays the same."""
        step = self._captured_step(evaluate, version)

        def replay(n):
            for _ in range(n):
                step.replay()
            self.waves_run += n
            self.evals_run += n
            self.rows_evaluated += n * self.rows_active

        self.wave()
        if not self._adaptive:
            replay(self.waves_per_search - 1)
            return
        while True:
            st = self.status()
            done = (st[:, 0] >= self.num_sims) | (st[:, 2] != 0)
            if bool(done.all()):
                return
            remaining = int((self.num_sims - st[:, 0])[~done].max())
            replay(max(1, -(-remaining // self.leaves_per_wave)))

    def wave_timing(self, enable: bool) -> None:
        """Bracket every (un-captured) wave launch with CUDA events inside the library call (see ifx_mcts_wave_timing)."""
        self.be.check(self.be.dll.ifx_mcts_wave_timing(self._h, 1 if enable else 0))

    def wave_time(self):
        """-> (summed wave-kernel milliseconds, launches) since the last read-out; synchronises on the events."""
        ms, n = C.c_double(0.0), C.c_int32(0)
        self.be.check(self.be.dll.ifx_mcts_wave_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- results --------------------------------------------------------------------------------
    def root_stats(self, want_q: bool = False) -> dict:
        g = self.n_games
        out = {"counts": self.be.empty((g, 343), torch.int32), "ns": self.be.empty((g,), torch.int32)}
        if want_q:
            out["q32"] = self.be.empty((g, 343), torch.float32)
            out["q64"] = self.be.empty((g, 343), torch.float64)
            out["qkind"] = self.be.empty((g, 343), torch.uint8)
        p = lambda k: out[k].data_ptr() if k in out else None  # noqa: E731
        self.be.check(self.be.dll.ifx_mcts_root_stats(self._h, p("counts"), p("q32"), p("q64"), p("qkind"), p("ns"),
                                                      self.be.stream()))
        return out

    def sample_actions(self, temp_threshold: int) -> torch.Tensor:
        a = self.be.empty((self.n_games,), torch.int32)
        self.be.check(self.be.dll.ifx_mcts_sample_actions(self._h, int(temp_threshold), a.data_ptr(), self.be.stream()))
        return a

    def advance(self, actions: torch.Tensor) -> None:
        assert actions.dtype == torch.int32 and actions.shape == (self.n_games,)
        self.be.check(self.be.dll.ifx_mcts_advance(self._h, actions.data_ptr(), self.be.stream()))
        self._keep_a = actions

    def restart_finished(self, tag_stride: int, count: torch.Tensor | None = None) -> None:
        """New episode (start position, empty tree, RNG stream id + tag_stride) for every game that is over."""
        self.be.check(self.be.dll.ifx_mcts_restart_finished(self._h, int(tag_stride), None if count is None else count.data_ptr(),
                                                            self.be.stream()))

    def status(self) -> torch.Tensor:
        """int32 [G,8]: sims_done, pending, fault bits, move index, root outcome, live nodes, free nodes, path length."""
        s = self.be.empty((self.n_games, 8), torch.int32)
        self.be.check(self.be.dll.ifx_mcts_status(self._h, s.data_ptr(), self.be.stream()))
        return s

    def raise_on_faults(self) -> None:
        f = self.status()[:, 2]
        if bool((f != 0).any()):
            bad = torch.nonzero(f).flatten()[:8].tolist()
            bits = int(f[bad[0]])
            raise _capi.IfxError(f"MCTS fault in games {bad}: {_capi.fault_text(bits)}")

    def roots(self) -> torch.Tensor:
        r = self.be.empty((self.n_games, 64), torch.uint8)
        self.be.check(self.be.dll.ifx_mcts_get_roots(self._h, r.data_ptr(), self.be.stream()))
        return r

    def counters(self) -> dict:
        c = self.be.empty((4,), torch.int64)
        self.be.check(self.be.dll.ifx_mcts_counters(self._h, c.data_ptr(), self.be.stream()))
        v, e, t, s = c.tolist()
        return {"node_visits": v, "expansions": e, "terminals": t, "sims": s}

    # ---- inspection (tests) -----------------------------------------------------------------------
    def table_views(self) -> dict:
        """numpy views of the SoA pools (copies the workspace to the host)."""
        ws = self.workspace.cpu().numpy()
        L, c = self.layout, self.cfg
        g, n, h = c.n_games, c.nodes_per_game, c.hash_per_game

        def view(off, dtype, shape):
            count = int(np.prod(shape))
            return np.frombuffer(ws, dtype=dtype, count=count, offset=off).reshape(shape)
        rows = view(L.rows, np.uint32, (g, n, 3, 352))
        q64c = view(L.q64, np.float64, (g, n, 344))
        keys = view(L.node_key, np.uint64, (g, n, 2))
        # rows are compact (entry i = the i-th valid action of the node, the valid set being a function of the node
        # key); expand them to dense [343] rows: invalid actions read P = 0, Nraw = 0xFFFFFFFF, Q = 0
        P = np.zeros((g, n, 343), np.float32)
        N = np.full((g, n, 343), 0xFFFFFFFF, np.uint32)
        Q32 = np.zeros((g, n, 343), np.float32)
        Q64 = np.zeros((g, n, 343), np.float64)
        board = (1 << 49) - 1
        for gi in range(g):
            for slot in np.nonzero(keys[gi, :, 0] != np.uint64(2 ** 64 - 1))[0]:
                k0, k1 = int(keys[gi, slot, 0]), int(keys[gi, slot, 1])
                own, opp = k0 & board, k1 & board
                spawn = (~(own | opp) & board) if (k1 >> 49) & 1 else 0
                valid = [m * 49 + c for m in range(6) for c in range(49) if (own >> c) & 1]
                valid += [6 * 49 + c for c in range(49) if (spawn >> c) & 1]
                idx = np.asarray(valid, dtype=np.int64)
                k = len(idx)
                P[gi, slot, idx] = rows[gi, slot, 0, :k].view(np.float32)
                N[gi, slot, idx] = rows[gi, slot, 1, :k]
                Q32[gi, slot, idx] = rows[gi, slot, 2, :k].view(np.float32)
                Q64[gi, slot, idx] = q64c[gi, slot, :k]
        return {"P": P, "Nraw": N, "Q32": Q32, "Q64": Q64, "key": keys,
                "ns": view(L.node_head, np.int32, (g, n, 4))[:, :, 0]}

"""``BatchedArena``: ``Arena.playGames`` (Arena.py:90-142) for an MCTS player against the Random or Greedy baseline
with ALL games running concurrently on the device -- what ``Coach.learn`` does every fifth iteration
(Coach.py:155-165) and what ``pit.py`` does.

Player 1 (RED) is the net-guided ``MCTSPlayer`` (arg-max of the root visit counts, ties broken uniformly,
InflexionPlayers.py:86-88 with MCTS.py:51-56; one fresh tree per game, kept across that player's moves);
player 2 (BLUE) is ``RandomPlayer`` or ``GreedyPlayer`` as a rules kernel.  The reference's colour split is kept:
of ``num`` games, game i has RED moving first iff ``i <= num // 2`` (Arena.py:126-129).
"""
from __future__ import annotations

import numpy as np
import torch

from . import rules
from .backend import Backend, default_backend
from .mcts_batched import BatchedMCTS
from .state import STATE_DTYPE, pack_states


class BatchedArena:
    def __init__(self, evaluate, num_sims: int, cpuct=1.0, opponent: str = "random", seed: int = 0,
                 max_turns: int = 343, backend: Backend | None = None, fused_head: bool = False, record: bool = False,
                 leaves_per_wave: int = 1, capture: bool = False):
        if opponent not in ("random", "greedy"):
            raise ValueError("opponent must be 'random' or 'greedy'")
        self.be = backend or default_backend()
        self.evaluate, self.num_sims, self.cpuct = evaluate, int(num_sims), cpuct
        self.opponent, self.seed, self.max_turns, self.fused_head = opponent, int(seed), int(max_turns), fused_head
        # record: keep every game's move sequence and result of the last playGames call (this rank's games):
        # ``last_actions`` int32 [plies, games] (-1 once a game is over), ``last_results`` (+1 RED won, -1 BLUE, 0 draw)
        self.record, self.last_actions, self.last_results = bool(record), None, None
        # leaves_per_wave > 1: the MCTS player searches leaf-parallel with virtual loss (performance mode: K times larger
        # net batches for the few games of an Arena; its moves then differ from the reference's sequential search)
        self.leaves_per_wave = int(leaves_per_wave)
        # capture (CUDA): the MCTS player's (net, wave) step is replayed from a CUDA graph (BatchedMCTS.search_captured) --
        # same moves, one launch per step; ``evaluate`` must then be capturable (device tensors only, no host sync)
        self.capture = bool(capture) and self.be.is_cuda

    def _opponent_actions(self, states, ply):
        if self.opponent == "random":
            return rules.random_actions(states, self.seed ^ 0x5EED, ply, backend=self.be)
        return rules.greedy_actions(states, backend=self.be)

    def playGames(self, num: int, shard: bool = True):
        """-> (games won by player 1 / RED / the MCTS player, games won by player 2 / BLUE, draws).

        Under torch.distributed (``shard``) the ``num`` games are split into contiguous blocks, one per rank --
        game i keeps its colour and its RNG stream whichever rank plays it -- and only the three counters are
        all-reduced; every rank returns the global result."""
        if not (isinstance(num, int) and num >= 2):
            raise ValueError("num must be an int >= 2")
        import torch.distributed as dist
        world = dist.get_world_size() if shard and dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        per = -(-num // world)
        lo, hi = min(num, rank * per), min(num, (rank + 1) * per)
        total, num = num, hi - lo
        if num == 0:
            result = torch.zeros(3, dtype=torch.int64, device=self.be.device)
        else:
            result = torch.tensor(self._play(total, lo, hi), dtype=torch.int64, device=self.be.device)
        if world > 1:
            dist.all_reduce(result)
        return tuple(int(x) for x in result.tolist())

    def _play(self, total: int, lo: int, hi: int):
        num = hi - lo
        red_first = np.arange(lo, hi) <= total // 2                             # Arena.py:126-129
        starts = pack_states(np.zeros((num, 49), dtype=np.int64), np.where(red_first, 1, -1), 0, 0, self.max_turns,
                             tags=np.arange(lo, hi))
        states = self.be.states_tensor(starts)
        mcts = BatchedMCTS(num, self.num_sims, self.cpuct, seed=self.seed, backend=self.be, fused_head=self.fused_head,
                           leaves_per_wave=self.leaves_per_wave)
        started = False
        played = []
        for ply in range(self.max_turns + 2):
            outcome = states[:, 45]
            if bool((outcome != 0).all()):
                break
            red_to_move = states[:, 44] == 0
            # the MCTS player searches the games where RED is to move; the others keep their state this half-ply
            if bool((red_to_move & (outcome == 0)).any()):
                # games where it is not RED's turn are handed to the search as finished (never searched)
                roots = states.clone()
                roots[:, 45] = torch.where(red_to_move, outcome, torch.full_like(outcome, 1))
                if not started:
                    mcts.reset(roots)
                    started = True
                else:
                    mcts.set_roots(roots)
                mcts.search(self.evaluate, pack=True, capture=self.capture)      # the net sees only the games RED moves in
                acts_red = mcts.sample_actions(0)                                # temp = 0: arg-max, ties uniformly
            else:
                acts_red = torch.full((num,), -1, dtype=torch.int32, device=self.be.device)
            acts_blue = self._opponent_actions(states, ply)
            acts = torch.where(red_to_move, acts_red, acts_blue).to(torch.int32).contiguous()
            _, nxt, status = rules.mask_step(states, acts, want_masks=False, backend=self.be)
            live = (outcome == 0)
            if self.record:
                played.append(torch.where(live, acts, torch.full_like(acts, -1)))
            if bool((status[live] != 0).any()):
                raise RuntimeError("a player produced an invalid action")       # Arena.py:65-68
            states = nxt
        mcts.raise_on_faults()
        final = states.cpu().numpy().view(STATE_DTYPE).reshape(-1)
        if (final["outcome"] == 0).any():
            raise RuntimeError("games did not finish")
        # outcome is relative to the side to move: WON -> that side won, LOST -> the other side won
        mover_red = final["player"] == 0
        red_won = ((final["outcome"] == 2) & mover_red) | ((final["outcome"] == 3) & ~mover_red)
        blue_won = ((final["outcome"] == 2) & ~mover_red) | ((final["outcome"] == 3) & mover_red)
        if self.record:
            self.last_actions = torch.stack(played).cpu().numpy() if played else np.zeros((0, num), np.int32)
            self.last_results = red_won.astype(np.int64) - blue_won.astype(np.int64)
        return int(red_won.sum()), int(blue_won.sum()), int((final["outcome"] == 1).sum())

"""Players with the reference's protocol (inflexion/InflexionPlayers.py:10-91): ``play(game) -> int`` and
``reset() -> Player``.  ``MCTSPlayer`` is the hot-path caller; Random and Greedy are the Arena baselines."""
from __future__ import annotations

import uuid

import numpy as np

from .. import rules
from ..Game import Game
from ..MCTS import MCTS
from .InflexionGame import InflexionGame, Move  # noqa: F401  (re-exported like the reference module)


class Player:
    """Protocol of the reference (InflexionPlayers.py:10-21): ``play(game) -> action``, ``reset() -> fresh player``,
    hashable by a per-object id (Arena keys its score table by player)."""

    def __init__(self):
        self.id = uuid.uuid4()

    def __hash__(self):
        return hash(self.id)

    def play(self, game: Game) -> int:
        raise NotImplementedError(f"{type(self).__name__}.play")

    def reset(self) -> "Player":
        return type(self)()                                  # stateless players: a new object of the same kind


class RandomPlayer(Player):
    """Uniform over the valid actions, one ``np.random.choice`` per move (InflexionPlayers.py:24-34)."""

    def play(self, game: Game) -> int:
        valid = np.flatnonzero(game.valid_actions_mask() == 1)
        return int(np.random.choice(valid))


class GreedyPlayer(Player):
    """One-ply look-ahead on ``score()`` (InflexionPlayers.py:61-77): all valid moves are stepped in ONE
    kernel call; ties go to the highest action index (the reference sorts (score, action) descending)."""

    def play(self, game: Game) -> int:
        valid = np.flatnonzero(game.valid_actions_mask())
        states = np.repeat(game._state, len(valid))
        _, nxt, status = rules.mask_step_host(states, valid.astype(np.int32), want_masks=False)
        assert not status.any()
        # score of the child from the mover's side = -(score for the child's side to move)
        scores = -rules.info_host(nxt)[:, 0]
        best = scores.max()
        return int(valid[np.flatnonzero(scores == best)[-1]])


class HumanPlayer(Player):
    """Reads "r q m" from stdin (InflexionPlayers.py:37-58)."""

    def play(self, game: Game) -> int:
        valid = game.valid_actions_mask()
        while True:
            print("Enter move as 3 integer: r q m")
            print("where m is in " + " ".join(f"{m.name}: {m.num}" for m in Move))
            r, q, m = (int(x) for x in input(">>>").split(" "))
            a = game.move_to_action((Move.from_num(m), r, q))
            if valid[a]:
                return a


class MCTSPlayer(Player):
    def __init__(self, mcts: MCTS):
        super().__init__()
        if not isinstance(mcts, MCTS):
            raise TypeError("MCTSPlayer needs an MCTS")
        self.mcts = mcts

    def play(self, game: Game) -> int:
        return int(np.argmax(self.mcts.getActionProb(game, temp=0)))      # InflexionPlayers.py:86-88

    def reset(self):
        return MCTSPlayer(self.mcts.reset())

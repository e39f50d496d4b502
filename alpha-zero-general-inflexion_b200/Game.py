"""Abstract two-player game interface (reference: Game.py:1-181).

A ``Game`` object is one position: board, side to move, turn counter and outcome.  Moves return new objects.
Concrete games implement the methods below; ``inflexion.InflexionGame`` is the device-backed implementation.
Every method names the reference lines whose contract it keeps.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .flags import GameOutcome, PlayerColour


def _abstract(name: str):
    return NotImplementedError(f"Game.{name} is implemented by the concrete game")


class Game:
    def __init__(self, board_shape: tuple, policy_shape: tuple, first_mover: PlayerColour):
        """Game.py:14-25.  ``board_shape`` = (m, n), ``policy_shape`` = (k, m, n) with k move types per cell."""
        if len(board_shape) != 2 or len(policy_shape) != 3 or not isinstance(first_mover, PlayerColour):
            raise TypeError("Game(board_shape=(m, n), policy_shape=(k, m, n), first_mover=PlayerColour)")
        self._board_shape = tuple(int(x) for x in board_shape)
        self._policy_shape = tuple(int(x) for x in policy_shape)
        self._firstMover = first_mover

    # -- shapes ---------------------------------------------------------------------------------
    @property
    def board_shape(self) -> tuple[int, int]:
        """(m, n) of the board (Game.py:27-33)."""
        return self._board_shape

    @property
    def policy_shape(self) -> tuple[int, int, int]:
        """(k, m, n) of a policy array (Game.py:35-43)."""
        return self._policy_shape

    @property
    def max_actions(self) -> int:
        """k * m * n: the length of a policy vector and of ``valid_actions_mask`` (Game.py:74-80)."""
        return int(np.prod(self._policy_shape))

    # -- position state (concrete classes store it) -------------------------------------------------
    @property
    def player(self) -> PlayerColour:
        """Side to move (Game.py:45-47)."""
        raise _abstract("player")

    @player.setter
    def player(self, player: PlayerColour):
        """Changing the side to move re-expresses the outcome from the new side's point of view (Game.py:49-62)."""
        raise _abstract("player")

    @property
    def outcome(self) -> GameOutcome:
        """ONGOING, or WON / LOST / DRAW from the point of view of the side to move (Game.py:64-72)."""
        raise _abstract("outcome")

    # -- rules ------------------------------------------------------------------------------------
    def restarted(self) -> "Game":
        """A fresh game with the same construction parameters (Game.py:82-87)."""
        raise _abstract("restarted")

    def to_next_state(self, action: int) -> "Game":
        """A NEW game object with ``action`` applied: turn + 1, side to move flipped, outcome updated; the receiver
        is never modified (Game.py:89-101)."""
        raise _abstract("to_next_state")

    def to_planes(self) -> np.ndarray:
        """The network input: feature planes plus constant planes for hidden state (Game.py:103-111).  Its bytes
        are also the reference's MCTS node key (MCTS.py:48,83)."""
        raise _abstract("to_planes")

    def valid_actions_mask(self) -> np.ndarray:
        """0/1 vector of length ``max_actions`` (Game.py:113-120)."""
        raise _abstract("valid_actions_mask")

    def symmetries(self, board_like: np.ndarray) -> list[np.ndarray]:
        """All isomorphic forms of an array whose last two axes are the board (Game.py:122-134)."""
        raise _abstract("symmetries")

    def random_symmetry(self, board_like: np.ndarray) -> np.ndarray:
        """One random isomorphic form; MCTS feeds the net through it (Game.py:136-148, MCTS.py:91-92)."""
        raise _abstract("random_symmetry")

    def score(self) -> int:
        """Heuristic score for the side to move (Game.py:150-155)."""
        raise _abstract("score")

    def move_to_action(self, move: tuple | list) -> int:
        """(move type, row, col) -> action index (Game.py:157-166)."""
        raise _abstract("move_to_action")

    def action_to_move(self, action: int) -> tuple[Any, int, int]:
        """action index -> (move type, row, col) (Game.py:168-177)."""
        raise _abstract("action_to_move")

    def render(self):
        """Print the board (Game.py:179-181)."""
        raise _abstract("render")

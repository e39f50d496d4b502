"""Abstract two-player game interface (reference: Game.py:1-181).

A ``Game`` object is one position: board, side to move, turn counter and outcome.  Moves return new objects.
Concrete games implement the methods below; ``inflexion.InflexionGame`` is the device-backed implementation.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .flags import GameOutcome, PlayerColour


class Game:
    def __init__(self, board_shape: tuple, policy_shape: tuple, first_mover: PlayerColour):
        if len(board_shape) != 2 or len(policy_shape) != 3 or not isinstance(first_mover, PlayerColour):
            raise TypeError("Game(board_shape=(m, n), policy_shape=(k, m, n), first_mover=PlayerColour)")
        self._board_shape = tuple(int(x) for x in board_shape)
        self._policy_shape = tuple(int(x) for x in policy_shape)
        self._firstMover = first_mover

    # -- shapes ---------------------------------------------------------------------------------
    @property
    def board_shape(self) -> tuple[int, int]:
        return self._board_shape

    @property
    def policy_shape(self) -> tuple[int, int, int]:
        return self._policy_shape

    @property
    def max_actions(self) -> int:
        return int(np.prod(self._policy_shape))

    # -- position state (concrete classes store it) -------------------------------------------------
    @property
    def player(self) -> PlayerColour:
        raise NotImplementedError

    @player.setter
    def player(self, player: PlayerColour):
        """Changing the side to move re-expresses the outcome from the new side's point of view."""
        raise NotImplementedError

    @property
    def outcome(self) -> GameOutcome:
        """ONGOING, or WON / LOST / DRAW from the point of view of the side to move."""
        raise NotImplementedError

    # -- rules ------------------------------------------------------------------------------------
    def restarted(self) -> "Game":
        raise NotImplementedError

    def to_next_state(self, action: int) -> "Game":
        raise NotImplementedError

    def to_planes(self) -> np.ndarray:
        raise NotImplementedError

    def valid_actions_mask(self) -> np.ndarray:
        raise NotImplementedError

    def symmetries(self, board_like: np.ndarray) -> list[np.ndarray]:
        raise NotImplementedError

    def random_symmetry(self, board_like: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def score(self) -> int:
        raise NotImplementedError

    def move_to_action(self, move: tuple | list) -> int:
        raise NotImplementedError

    def action_to_move(self, action: int) -> tuple[Any, int, int]:
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

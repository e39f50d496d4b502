"""``InflexionGame``: the reference's game object API (inflexion/InflexionGame.py:39-348) over the CUDA rules
engine.  One object = one packed 64-byte ``ifx_state``; every rule evaluation (move masks, stepping, planes,
symmetries, scores) is a call into libifx_b200 through the host-buffer C ABI.  Thousands of positions at once
go through ``inflexion_b200.rules`` / ``BatchedMCTS`` instead -- this class is the drop-in for code written
against the reference (Coach.executeEpisode, Arena, pit.py).

Only the 7x7 board is implemented (the kernels are specialised to it; main.py:34 / pit.py:41 use n=7).
"""
from __future__ import annotations

from enum import Enum

import numpy as np

from .. import rules
from ..flags import OUTCOME_FROM_CODE, OUTCOME_TO_CODE, GameOutcome, PlayerColour
from ..Game import Game
from ..state import ACTIONS, CELLS, N, pack_states, unpack_boards, unpack_masks


class Move(Enum):
    """Move types and their (dr, dq) directions (InflexionGame.py:14-36)."""
    SPREAD_R1 = 0, (1, 0)
    SPREAD_R2 = 1, (-1, 0)
    SPREAD_Q1 = 2, (0, 1)
    SPREAD_Q2 = 3, (0, -1)
    SPREAD_P1 = 4, (1, -1)
    SPREAD_P2 = 5, (-1, 1)
    SPAWN = 6, (0, 0)

    def __init__(self, num, direction):
        self.num = num
        self.direction = direction

    @classmethod
    def from_num(cls, num: int) -> "Move":
        for move in cls:
            if move.num == num:
                return move
        raise IndexError(f"Move number {num} is not valid.")

    @classmethod
    def all_spreads(cls) -> tuple["Move", ...]:
        return tuple(m for m in cls if m is not cls.SPAWN)


_AXES = "rqs"


class InflexionGame(Game):
    def __init__(self, n: int, first_mover: PlayerColour = PlayerColour.RED, curr_player: PlayerColour = None,
                 board: np.ndarray = None, curr_turn: int = 0, max_turns: int = 100, max_power: int = 6):
        super().__init__(board_shape=(n, n), policy_shape=(7, n, n), first_mover=first_mover)
        if n != N:
            raise NotImplementedError("the CUDA rules engine is specialised to the 7x7 board (main.py:34)")
        if board is not None and np.asarray(board).shape != (n, n):
            raise ValueError("board must have shape (n, n)")
        self._n = n
        self._max_turns = int(max_turns)
        self._max_power = int(max_power)          # stored, unused -- as in the reference (InflexionGame.py:66,288)
        self._planes_shape = (4, n, n)
        self._max_power_at_spawn = 48
        player = first_mover if curr_player is None else curr_player
        cells = np.zeros(CELLS, dtype=np.int64) if board is None else np.asarray(board, dtype=np.int64).reshape(CELLS)
        self._state = pack_states(cells[None], player.num, int(curr_turn), 0, self._max_turns)

    # ---- position state ------------------------------------------------------------------------------
    @classmethod
    def _from_state(cls, proto: "InflexionGame", state: np.ndarray) -> "InflexionGame":
        g = cls.__new__(cls)
        g.__dict__.update(proto.__dict__)
        g._state = state
        return g

    @property
    def _board(self) -> np.ndarray:
        """int64 [7,7] signed stack powers (read by inflexion/utils.py:67-68 and by user code)."""
        return unpack_boards(self._state).reshape(self._n, self._n)

    @property
    def _curr_turn(self) -> int:                   # read by Coach.py:59 and Arena.py:52
        return int(self._state["turn"][0])

    @property
    def player(self) -> PlayerColour:
        return PlayerColour.BLUE if self._state["player"][0] else PlayerColour.RED

    @player.setter
    def player(self, player: PlayerColour):
        """Game.py:49-62: switching the side to move flips WON <-> LOST."""
        if player is self.player:
            return
        self._state = self._state.copy()
        self._state["player"] = 1 if player is PlayerColour.BLUE else 0
        self._state["outcome"] = OUTCOME_TO_CODE[self.outcome.opposite()]

    @property
    def outcome(self) -> GameOutcome:
        return OUTCOME_FROM_CODE[int(self._state["outcome"][0])]

    # ---- rules (each one is a kernel call) -------------------------------------------------------------
    def restarted(self) -> "InflexionGame":
        return InflexionGame(self._n, first_mover=self._firstMover, max_turns=self._max_turns, max_power=self._max_power)

    def to_next_state(self, action: int) -> "InflexionGame":
        """A new game with ``action`` played by the side to move (InflexionGame.py:75-82); the receiver is untouched."""
        _, nxt, status = rules.mask_step_host(self._state, np.array([action], dtype=np.int32), want_masks=False)
        if status[0]:
            raise ValueError(f"Invalid move: action {action}")      # InflexionGame.py:279,282,290-291
        return self._from_state(self, nxt)

    def execute_move(self, move: tuple):
        """In-place form (InflexionGame.py:273-310): ``move`` = (Move, r, q)."""
        self._state = self.to_next_state(self.move_to_action(move))._state

    def to_planes(self) -> np.ndarray:
        """int64 [4,7,7]: own stacks, opponent stacks, turn counter, spawn-allowed (InflexionGame.py:84-91)."""
        return rules.planes_host(self._state, dtype=np.int64)[0]

    def valid_actions_mask(self) -> np.ndarray:
        """int64 [343], a = move*49 + r*7 + q (InflexionGame.py:93-100)."""
        masks, _, _ = rules.mask_step_host(self._state, None)
        return unpack_masks(masks)[0].astype(np.int64)

    def _board_like(self, x) -> np.ndarray:
        x = np.asarray(x)
        if x.ndim != 3 or x.shape[-2:] != self.board_shape:
            raise ValueError("expected a [channels, 7, 7] array (planes or policy)")
        return x

    def _gather(self, x: np.ndarray, fn):
        """Run a device gather on any dtype: 4/8-byte dtypes as they are, others through int64/float64."""
        if x.dtype.itemsize in (4, 8):
            return fn(np.ascontiguousarray(x)[None])[0]
        wide = np.float64 if x.dtype.kind == "f" else np.int64
        return fn(np.ascontiguousarray(x.astype(wide))[None])[0].astype(x.dtype)

    def symmetries(self, board_like: np.ndarray) -> list[np.ndarray]:
        """The 36 isomorphic forms in the reference's order (InflexionGame.py:102-113)."""
        x = self._board_like(board_like)
        return list(self._gather(x, rules.symmetries_host))

    def rotate(self, board_like: np.ndarray, k: int = 1) -> np.ndarray:
        """Rotation by 60*k degrees about cell (0,0) (InflexionGame.py:124-168)."""
        x = self._board_like(board_like)
        return self._gather(x, lambda a: rules.transform_host(a, [int(k) % 6]))

    def translate(self, board_like: np.ndarray, shift: int, axis: str) -> np.ndarray:
        """Torus shift along the r, q or s axis (InflexionGame.py:170-196)."""
        if axis not in _AXES or len(axis) != 1:
            raise ValueError("axis must be 'r', 'q' or 's'")
        x = self._board_like(board_like)
        return self._gather(x, lambda a: rules.transform_host(a, [6 * (int(shift) % N) + 42 * _AXES.index(axis)]))

    def random_symmetry(self, board_like: np.ndarray) -> np.ndarray:
        """A random rotation then translation; draws from numpy's global RNG in the reference's order
        (InflexionGame.py:115-122), so a seeded run picks the same symmetry as the reference."""
        x = self._board_like(board_like)
        k = int(np.random.randint(0, 6))
        shift = int(np.random.randint(0, self._n))
        axis = _AXES.index(np.random.choice(["r", "q", "s"]))
        return self._gather(x, lambda a: rules.transform_host(a, [k + 6 * shift + 42 * axis]))

    def _info(self):
        return rules.info_host(self._state)[0]

    def score(self) -> int:
        """#own stacks - #opponent stacks for the side to move (InflexionGame.py:198-199)."""
        return int(self._info()[0])

    def piece_count_diff(self, player: PlayerColour) -> int:
        s = int(self._info()[0])
        return s if player is self.player else -s

    def power_diff(self, player: PlayerColour) -> int:
        d = int(self._info()[1])
        return d if player is self.player else -d

    def total_power(self) -> int:
        return int(self._info()[2])

    # ---- action <-> move (InflexionGame.py:201-210) --------------------------------------------------
    def move_to_action(self, move: tuple | list) -> int:
        move_type, r, q = move
        if not (0 <= r < self._n and 0 <= q < self._n and isinstance(move_type, Move)):
            raise ValueError(f"bad move {move}")
        return int(move_type.num * CELLS + r * self._n + q)

    def action_to_move(self, action: int) -> tuple[Move, int, int]:
        if not 0 <= action < ACTIONS:
            raise ValueError(f"bad action {action}")
        m, cell = divmod(int(action), CELLS)
        return Move.from_num(m), cell // self._n, cell % self._n

    # ---- display --------------------------------------------------------------------------------------
    def render(self, ansi=False):
        """Print the board as a hexagon in axial coordinates (same layout as InflexionGame.py:212-271)."""
        b = self._board
        dim = self._n
        lines = []
        for row in range(2 * dim - 1):
            cells = []
            for col in range(dim - abs(row - (dim - 1))):
                r = max((dim - 1) - row, 0) + col
                q = max(row - (dim - 1), 0) + col
                v = int(b[r, q])
                if v == 0:
                    text = " .. "
                else:
                    token = PlayerColour.from_piece(v).token
                    text = f"{token}{abs(v)}".center(4)
                    if ansi:
                        text = self.apply_ansi(text, color=token, bold=False)
                cells.append(text)
            lines.append("    " * abs((dim - 1) - row) + "    ".join(cells) + "    ")
        print("\n".join(lines) + "\n")

    @staticmethod
    def apply_ansi(str, bold=True, color=None):
        codes = {PlayerColour.RED.token: "\033[31m", PlayerColour.BLUE.token: "\033[34m"}
        return f"{chr(27) + '[1m' if bold else ''}{codes.get(color, '')}{str}\033[0m"

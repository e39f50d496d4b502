"""``PlayerColour`` and ``GameOutcome``: the two enums every caller of the path exchanges.

Member names, member values and the public helpers (``num``, ``token``, ``from_piece``, ``opponent``, ``owns``,
``opposite``) are those of the reference (flags.py:1-44) because game objects, pickled examples and callers depend on
them; the bodies are table look-ups here.  ``OUTCOME_FROM_CODE`` / ``OUTCOME_TO_CODE`` translate to and from the
device's outcome codes (include/ifx_b200.h).
"""
from enum import Enum


class PlayerColour(Enum):
    """Side to move / stack owner.  ``num`` is the sign of the side's stacks in ``InflexionGame._board``."""
    RED = (1, "R")
    BLUE = (-1, "B")

    def __init__(self, sign: int, letter: str):
        self.num, self.token = sign, letter

    @classmethod
    def from_num(cls, num: int) -> "PlayerColour":
        return cls.BLUE if num < 0 else cls.RED

    @classmethod
    def from_piece(cls, piece) -> "PlayerColour":
        if piece == 0:
            raise IndexError(f"No player owns piece {piece}")
        return cls.from_num(piece)

    @property
    def opponent(self) -> "PlayerColour":
        return _OTHER_SIDE[self]

    def owns(self, piece):
        """True where a stack (int, float or ndarray of signed powers) belongs to this side."""
        return (piece > 0) if self is PlayerColour.RED else (piece < 0)


class GameOutcome(Enum):
    """Result seen by the side to move; ``value`` is what MCTS backs up (a draw is worth 1e-4, flags.py:34)."""
    ONGOING = 0
    DRAW = 1e-4
    WON = 1
    LOST = -1

    def opposite(self) -> "GameOutcome":
        """The same result seen by the other side."""
        return _OTHER_VIEW.get(self, self)


_OTHER_SIDE = {PlayerColour.RED: PlayerColour.BLUE, PlayerColour.BLUE: PlayerColour.RED}
_OTHER_VIEW = {GameOutcome.WON: GameOutcome.LOST, GameOutcome.LOST: GameOutcome.WON}

# device outcome codes (include/ifx_b200.h) <-> enum
OUTCOME_FROM_CODE = (GameOutcome.ONGOING, GameOutcome.DRAW, GameOutcome.WON, GameOutcome.LOST)
OUTCOME_TO_CODE = {o: i for i, o in enumerate(OUTCOME_FROM_CODE)}

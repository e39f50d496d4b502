"""``PlayerColour`` and ``GameOutcome`` with the reference's members, values and helpers (flags.py:1-44)."""
from enum import Enum


class PlayerColour(Enum):
    RED = 1, "R"
    BLUE = -1, "B"

    def __init__(self, num, token):
        self.num = num            # sign of this player's stacks on the board
        self.token = token

    @classmethod
    def from_num(cls, num: int) -> "PlayerColour":
        return cls.RED if num > 0 else cls.BLUE

    @classmethod
    def from_piece(cls, piece):
        if piece > 0:
            return cls.RED
        if piece < 0:
            return cls.BLUE
        raise IndexError(f"No player owns piece {piece}")

    @property
    def opponent(self) -> "PlayerColour":
        return PlayerColour.BLUE if self is PlayerColour.RED else PlayerColour.RED

    def owns(self, piece):
        """Works on ints and on numpy arrays: True where the stack belongs to this player."""
        return piece * self.num > 0


class GameOutcome(Enum):
    ONGOING = 0
    DRAW = 1e-4      # a draw is worth a little (flags.py:34); MCTS backs this value up
    WON = 1
    LOST = -1

    def opposite(self) -> "GameOutcome":
        if self is GameOutcome.WON:
            return GameOutcome.LOST
        if self is GameOutcome.LOST:
            return GameOutcome.WON
        return self


# device outcome codes (include/ifx_b200.h) <-> enum
OUTCOME_FROM_CODE = (GameOutcome.ONGOING, GameOutcome.DRAW, GameOutcome.WON, GameOutcome.LOST)
OUTCOME_TO_CODE = {o: i for i, o in enumerate(OUTCOME_FROM_CODE)}

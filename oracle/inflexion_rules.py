"""CPU oracle for the Inflexion rules engine.  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy restatement of the rules half of the hot path of
Meatssauce/alpha-zero-general-inflexion (``inflexion/InflexionGame.py``).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it -- never the product package.

Parity pin: the reference ships no golden vectors for this path (its single
test, ``test_all_games.py``, is an un-asserted smoke test and is broken, see
SURVEY.md section 4), so the pin is the LIVE reference: ``tests/golden/
make_golden.py`` imports ``/root/reference`` in the build container, replays
random games through the reference's ``InflexionGame`` and freezes the
transitions under ``tests/golden/``; ``tests/test_oracle_rules.py`` checks
this oracle against those files (and against the live reference when it is
mounted).

Board convention (``InflexionGame.py:57-69``): ``board[r, q]`` is a signed stack
power in [-6, 6]; the sign is the owner (RED = +1, BLUE = -1, ``flags.py:6-8``).
Action index ``a = m*49 + r*7 + q`` with move number ``m`` in 0..6
(``InflexionGame.py:14-21, 201-210``).
"""
from __future__ import annotations

import numpy as np

N = 7                      # board edge, main.py:34
CELLS = N * N              # 49
MOVES = 7                  # 6 spreads + spawn, InflexionGame.py:14-21
ACTIONS = MOVES * CELLS    # 343
SPAWN = 6
MAX_POWER_AT_SPAWN = 48    # InflexionGame.py:68
# (dr, dq) per move number, InflexionGame.py:15-21
DIRECTIONS = ((1, 0), (-1, 0), (0, 1), (0, -1), (1, -1), (-1, 1), (0, 0))

# outcome codes shared with the packed device state; values follow flags.py:32-36
ONGOING, DRAW, WON, LOST = 0, 1, 2, 3
OUTCOME_VALUE = {ONGOING: 0, DRAW: 1e-4, WON: 1, LOST: -1}   # WON/LOST are Python ints there


def flip_outcome(o: int) -> int:
    """flags.py:38-44 -- WON <-> LOST, DRAW and ONGOING unchanged."""
    return LOST if o == WON else WON if o == LOST else o


class OracleGame:
    """Scalar restatement of ``InflexionGame`` (state + rules), used by the MCTS oracle.

    ``player`` is +1 (RED) or -1 (BLUE); ``outcome`` is relative to the side to
    move (``Game.py:64-72``).
    """

    __slots__ = ("board", "player", "turn", "outcome", "max_turns")

    def __init__(self, board=None, player=1, turn=0, outcome=ONGOING, max_turns=343):
        self.board = np.zeros((N, N), dtype=np.int64) if board is None else board
        self.player = player
        self.turn = turn
        self.outcome = outcome
        self.max_turns = max_turns

    def copy(self) -> "OracleGame":
        return OracleGame(self.board.copy(), self.player, self.turn, self.outcome, self.max_turns)

    def restarted(self) -> "OracleGame":           # InflexionGame.py:72-73
        return OracleGame(max_turns=self.max_turns)

    def total_power(self) -> int:                  # InflexionGame.py:347-348
        return int(np.abs(self.board).sum())

    def can_spawn(self) -> bool:
        return self.total_power() <= MAX_POWER_AT_SPAWN

    def valid_actions_mask(self) -> np.ndarray:    # InflexionGame.py:93-100
        own = (self.board * self.player) > 0
        valid = np.zeros((MOVES, N, N), dtype=np.int64)
        valid[:SPAWN] = own
        if self.can_spawn():
            valid[SPAWN] = self.board == 0
        return valid.reshape(ACTIONS)

    def to_planes(self) -> np.ndarray:             # InflexionGame.py:84-91
        planes = np.empty((4, N, N), dtype=np.int64)
        planes[0] = (self.board * self.player) > 0
        planes[1] = (self.board * -self.player) > 0
        planes[2] = self.turn
        planes[3] = int(self.can_spawn())
        return planes

    def key(self) -> bytes:                        # MCTS.py:48,83
        return self.to_planes().tobytes()

    def apply(self, action: int) -> None:
        """In-place ``execute_move`` (``InflexionGame.py:273-310``)."""
        m, cell = divmod(int(action), CELLS)
        r, q = divmod(cell, N)
        b, p = self.board, self.player
        if m == SPAWN:
            if not self.can_spawn():               # InflexionGame.py:278,290-291
                raise ValueError("Invalid move")
            b[r, q] = p
        else:
            k = abs(int(b[r, q]))
            dr, dq = DIRECTIONS[m]
            for i in range(1, k + 1):              # k <= 6 < 7: cells distinct, source excluded
                rr, qq = (r + i * dr) % N, (q + i * dq) % N
                x = abs(int(b[rr, qq])) + 1
                b[rr, qq] = (0 if x > 6 else x) * p
            b[r, q] = 0
        # outcome for the mover, first match wins (InflexionGame.py:294-306)
        out = ONGOING
        if m != SPAWN and not ((b * -p) > 0).any():
            out = WON
        elif self.turn >= self.max_turns:
            diff = p * int(b.sum())
            out = WON if diff >= 2 else LOST if diff <= -2 else DRAW
        elif not b.any():
            out = DRAW
        self.turn += 1
        self.player = -p
        self.outcome = flip_outcome(out)           # Game.py:59-62 via InflexionGame.py:310

    def to_next_state(self, action: int) -> "OracleGame":   # InflexionGame.py:75-82
        nxt = self.copy()
        nxt.apply(action)
        return nxt

    def score(self) -> int:                        # InflexionGame.py:198-199, 319-323
        return int(((self.board * self.player) > 0).sum() - ((self.board * -self.player) > 0).sum())


# ----------------------------------------------------------------------------------------------
# Batched restatement (vectorised over S states) -- used to check the rules kernel at 1M states.
# ----------------------------------------------------------------------------------------------

def _dir_tables():
    """dest[m, cell, i-1] = flat index of (cell + i*direction[m]) mod 7 for i = 1..6."""
    dest = np.zeros((6, CELLS, 6), dtype=np.int64)
    for m in range(6):
        dr, dq = DIRECTIONS[m]
        for cell in range(CELLS):
            r, q = divmod(cell, N)
            for i in range(1, 7):
                dest[m, cell, i - 1] = ((r + i * dr) % N) * N + (q + i * dq) % N
    return dest


_DEST = _dir_tables()


def batch_valid_mask(boards: np.ndarray, players: np.ndarray) -> np.ndarray:
    """boards int8/int64 [S,49], players int [S] (+1/-1) -> uint8 [S,343] (InflexionGame.py:93-100)."""
    b = boards.reshape(len(boards), CELLS).astype(np.int64)
    own = (b * players[:, None]) > 0
    spawn_ok = np.abs(b).sum(axis=1) <= MAX_POWER_AT_SPAWN
    out = np.zeros((len(b), MOVES, CELLS), dtype=np.uint8)
    out[:, :SPAWN, :] = own[:, None, :]
    out[:, SPAWN, :] = (b == 0) & spawn_ok[:, None]
    return out.reshape(len(b), ACTIONS)


def batch_planes(boards, players, turns) -> np.ndarray:
    """-> int64 [S,4,7,7] (InflexionGame.py:84-91)."""
    b = boards.reshape(len(boards), CELLS).astype(np.int64)
    s = len(b)
    planes = np.empty((s, 4, CELLS), dtype=np.int64)
    planes[:, 0] = (b * players[:, None]) > 0
    planes[:, 1] = (b * -players[:, None]) > 0
    planes[:, 2] = np.asarray(turns, dtype=np.int64)[:, None]
    planes[:, 3] = (np.abs(b).sum(axis=1) <= MAX_POWER_AT_SPAWN)[:, None]
    return planes.reshape(s, 4, N, N)


def batch_step(boards, players, turns, actions, max_turns=343):
    """Apply one VALID action per state (InflexionGame.py:273-310).

    Returns (next_boards int64 [S,49], next_players, next_turns, outcomes) with outcomes
    relative to the NEW side to move (codes ONGOING/DRAW/WON/LOST above).
    """
    b = boards.reshape(len(boards), CELLS).astype(np.int64).copy()
    players = np.asarray(players, dtype=np.int64)
    turns = np.asarray(turns, dtype=np.int64)
    actions = np.asarray(actions, dtype=np.int64)
    s = len(b)
    idx = np.arange(s)
    m, cell = np.divmod(actions, CELLS)
    spawn = m == SPAWN
    # SPAWN
    b[idx[spawn], cell[spawn]] = players[spawn]
    # SPREAD
    sp = ~spawn
    k = np.abs(b[idx, cell]) * sp
    for i in range(1, 7):
        sel = np.nonzero(k >= i)[0]
        if len(sel) == 0:
            break
        d = _DEST[m[sel], cell[sel], i - 1]
        x = np.abs(b[sel, d]) + 1
        b[sel, d] = np.where(x > 6, 0, x) * players[sel]
    b[idx[sp], cell[sp]] = 0
    # outcome for the mover
    opp_cells = ((b * -players[:, None]) > 0).sum(axis=1)
    diff = players * b.sum(axis=1)
    out = np.full(s, ONGOING, dtype=np.int64)
    empty = ~b.any(axis=1)
    out[empty] = DRAW
    late = turns >= max_turns
    out[late] = np.where(diff[late] >= 2, WON, np.where(diff[late] <= -2, LOST, DRAW))
    wipe = sp & (opp_cells == 0)
    out[wipe] = WON
    # flip to the new side to move
    flipped = np.where(out == WON, LOST, np.where(out == LOST, WON, out))
    return b, -players, turns + 1, flipped


# ----------------------------------------------------------------------------------------------
# Symmetries (InflexionGame.py:102-196; index maps SURVEY.md appendix C)
# ----------------------------------------------------------------------------------------------

_SIGN_FLIP = {0: (1, 1, 1), 1: (-1, 1, 1), 2: (-1, 1, -1), 3: (-1, -1, -1), 4: (1, -1, -1), 5: (1, -1, 1)}


def rotate(x: np.ndarray, k: int = 1) -> np.ndarray:
    """InflexionGame.py:124-168: roll the (r, q, s) coordinate stack by k and flip signs."""
    k %= 6
    r, q = np.indices((N, N))
    s = (r + q) % N
    coords = np.roll(np.stack([r, q, s]), k, axis=0) * np.array(_SIGN_FLIP[k]).reshape(3, 1, 1)
    return x[:, coords[0], coords[1]].copy()       # negative indices wrap (== mod 7)


def translate(x: np.ndarray, shift: int, axis: str) -> np.ndarray:
    """InflexionGame.py:170-196."""
    if axis == "r":
        return np.roll(x, shift, axis=1)
    if axis == "q":
        return np.roll(x, shift, axis=2)
    if axis == "s":
        return np.roll(np.roll(x, shift, axis=2), -shift, axis=1)
    raise ValueError(axis)


def symmetries(x: np.ndarray) -> list[np.ndarray]:
    """InflexionGame.py:102-113: identity, 5 rotations, 5x6 r-translations of those = 36."""
    rotated = [rotate(x, k) for k in range(1, 6)]
    translated = [translate(rk, j, "r") for rk in rotated for j in range(1, N)]
    return [x.copy()] + rotated + translated


def random_symmetry(x: np.ndarray, k: int, shift: int, axis_idx: int) -> np.ndarray:
    """InflexionGame.py:115-122 with the three draws injected (k in [0,6), shift in [0,7), axis in 'rqs')."""
    return translate(rotate(x, k), shift, "rqs"[axis_idx])


def random_reachable_states(n_states: int, seed: int = 0, max_turns: int = 343):
    """cfg2 generator: uniform-random valid play from the start for a uniform number
    of plies in [0, max_turns], vectorised.  Returns (boards int8 [S,49], players, turns, actions)
    where ``actions`` holds one uniform-random valid action per final state."""
    rng = np.random.default_rng(seed)
    plies = rng.integers(0, max_turns + 1, size=n_states)
    b = np.zeros((n_states, CELLS), dtype=np.int64)
    players = np.ones(n_states, dtype=np.int64)
    turns = np.zeros(n_states, dtype=np.int64)

    def pick(bb, pp):
        mask = batch_valid_mask(bb, pp)
        cnt = mask.sum(axis=1)
        assert (cnt > 0).all()
        nth = (rng.random(len(bb)) * cnt).astype(np.int64)
        csum = np.cumsum(mask, axis=1)
        return (csum <= nth[:, None]).sum(axis=1)

    for t in range(max_turns):
        live = np.nonzero(plies > t)[0]
        if len(live) == 0:
            break
        a = pick(b[live], players[live])
        nb, npl, nt, out = batch_step(b[live], players[live], turns[live], a, max_turns)
        # a finished game (wipe-out) stays where it ended
        done = out != ONGOING
        keep = live[~done]
        b[keep], players[keep], turns[keep] = nb[~done], npl[~done], nt[~done]
        plies[live[done]] = t
    actions = pick(b, players)
    return b.astype(np.int8), players, turns, actions

"""``Arena``: pits two players against each other (reference: Arena.py:18-146).

Same constructor, ``playGame`` / ``playGames`` / ``getGameResults`` signatures and result conventions.  Games
run in this process one after another -- every move is already a kernel call, and the reference's
``torch.multiprocessing.Pool`` (Arena.py:125-133) exists only to spread Python work over cores.  The
reference's colour split (``i <= num // 2``, Arena.py:126-129) is kept.
"""
from __future__ import annotations

import logging
from collections import Counter
from itertools import cycle

from .flags import GameOutcome, PlayerColour

log = logging.getLogger(__name__)


class Arena:
    def __init__(self, player1, player2, game):
        self.player = {PlayerColour.RED: player1, PlayerColour.BLUE: player2}
        self.game = game

    def playGame(self, player1: PlayerColour, player2: PlayerColour, verbose=False):
        """One game; ``player1`` (a colour) moves first.  Returns (Counter of wins by colour, draws)."""
        if not (isinstance(player1, PlayerColour) and isinstance(player2, PlayerColour)):
            raise TypeError("playGame takes two PlayerColour values")
        game = self.game.restarted()
        game.player = player1
        movers = cycle([self.player[player1].reset(), self.player[player2].reset()])    # fresh search trees per game
        it = 0
        while game.outcome is GameOutcome.ONGOING:
            it += 1
            if verbose:
                print("Turn ", str(it), "Player ", str(game.player.name))
                game.render()
            action = next(movers).play(game)
            if game.valid_actions_mask()[action] == 0:
                log.error(f"Action {action} is not valid!")
                raise AssertionError(f"Action {action} is not valid!")                   # Arena.py:65-68
            game = game.to_next_state(action)
        game.player = player1                                                             # result from player1's side
        if verbose:
            print("Game over: Turn ", str(it), "Result ", str(game.outcome.name))
            game.render()
        wins, draws = Counter(), 0
        if game.outcome is GameOutcome.WON:
            wins[game.player] += 1
        elif game.outcome is GameOutcome.LOST:
            wins[game.player.opponent] += 1
        elif game.outcome is GameOutcome.DRAW:
            draws += 1
        else:
            raise ValueError(f"Unexpected game status: {game.outcome}")
        return wins, draws

    def getGameResults(self, args):
        player1, player2, verbose = args
        return player1, self.playGame(player1=player1, player2=player2, verbose=verbose)

    def playGames(self, num: int, verbose=False):
        """``num`` games; returns (won by player1, won by player2, draws)."""
        if not (isinstance(num, int) and num >= 2):
            raise ValueError("num must be an int >= 2")
        total, draws_total = Counter(), 0
        subtotal = num // 2
        red, blue = list(PlayerColour)
        for i in range(num):
            first, second = (red, blue) if i <= subtotal else (blue, red)
            _, (wins, draws) = self.getGameResults((first, second, verbose))
            total += wins
            draws_total += draws
        return total[red], total[blue], draws_total

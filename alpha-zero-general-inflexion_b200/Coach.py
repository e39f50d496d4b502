"""``Coach``: self-play + learning loop with the reference's interface (Coach.py:25-193).

``executeEpisode((game, mcts))`` is the reference's episode, move by move, over the drop-in classes
(one game, reference-identical examples).  ``learn()`` runs the same iteration structure -- self-play,
example history, training, checkpoint, periodic pit against Random/Greedy -- but plays the ``numEps``
self-play games of an iteration CONCURRENTLY on the device with ``SelfPlayEngine`` (and across ranks when
torch.distributed is initialised: games shard by rank, gradients all-reduce in ``NNetWrapper.train``).
Example files keep the reference's format: a pickle of ``list[deque[(planes int64[4,7,7], pi list[343], v)]]``.
"""
from __future__ import annotations

import logging
import os
import sys
from collections import deque
from pickle import Pickler, Unpickler
from random import shuffle

import numpy as np

from .Arena import Arena
from .flags import GameOutcome
from .Game import Game
from .inflexion.InflexionPlayers import GreedyPlayer, MCTSPlayer, RandomPlayer
from .MCTS import MCTS

log = logging.getLogger(__name__)


class Coach:
    def __init__(self, game: Game, nnet, args):
        self.game = game
        self.nnet = nnet
        self.args = args
        self.trainExamplesHistory = []      # one deque per iteration, newest last (Coach.py:37)
        self.skipFirstSelfPlay = False

    # ---- one game, object API (Coach.py:41-90) -----------------------------------------------------
    def executeEpisode(self, args):
        game, mcts = args
        if not (isinstance(game, Game) and isinstance(mcts, MCTS)):
            raise TypeError("executeEpisode((game, mcts))")
        assert game._curr_turn == 0 and game.outcome is GameOutcome.ONGOING
        policy_planes, board_stacks, players = [], [], []
        step = 0
        while True:
            step += 1
            temp = int(step < self.args.tempThreshold)
            pi = mcts.getActionProb(game, temp=temp)
            policy_planes += game.symmetries(pi.reshape(game.policy_shape))
            board_stacks += game.symmetries(game.to_planes())
            players += [game.player] * len(policy_planes)           # cumulative length: the reference's quirk (D1)
            action = int(np.random.choice(len(pi), p=pi))
            game = game.to_next_state(action)
            result = game.outcome
            if result is GameOutcome.ONGOING:
                continue
            return [(board, policy.ravel().tolist(), result.value if player == game.player else -result.value)
                    for board, policy, player in zip(board_stacks, policy_planes, players)]

    # ---- batched self-play for learn() -----------------------------------------------------------------
    def selfPlayBatch(self, num_games: int, seed: int = 0, game_id_base: int = 0):
        """``num_games`` episodes at once on the device -> list of (planes, pi list, v) in the reference's order."""
        from .selfplay import SelfPlayEngine
        eng = SelfPlayEngine(num_games, int(self.args.numMCTSSims), self.nnet.predict_batch, cpuct=self.args.cpuct,
                             temp_threshold=int(self.args.tempThreshold), seed=seed, max_turns=self.game._max_turns,
                             game_id_base=game_id_base)
        eng.play_episode()
        eng.mcts.raise_on_faults()
        planes, pi, v = eng.examples()
        return [(planes[i], pi[i].tolist(), float(v[i])) for i in range(len(v))]

    def learn(self):
        pit_interval = 5
        for i in range(1, self.args.numIters + 1):
            log.info(f"Starting Iter #{i} ...")
            if not self.skipFirstSelfPlay or i > 1:
                examples = deque([], maxlen=self.args.maxlenOfQueue)
                examples += self.selfPlayBatch(int(self.args.numEps), seed=int(np.random.randint(0, 2 ** 31 - 1)))
                self.trainExamplesHistory.append(examples)
            if len(self.trainExamplesHistory) > self.args.numItersForTrainExamplesHistory:
                log.warning(f"Removing the oldest entry in trainExamples. len(trainExamplesHistory) = "
                            f"{len(self.trainExamplesHistory)}")
                self.trainExamplesHistory.pop(0)
            self.saveTrainExamples(i - 1)
            train_examples = []
            for e in self.trainExamplesHistory:
                train_examples.extend(e)
            shuffle(train_examples)
            self.nnet.train(train_examples)
            self.nnet.save_checkpoint(folder=self.args.checkpoint, filename="temp.pth.tar")
            if i % pit_interval != 0:
                continue
            log.info("PITTING AGAINST BASELINES")
            if hasattr(self.nnet, "predict_batch"):
                # all arenaCompare games of a pairing at once on the device (same players, same colour split)
                from .arena_batched import BatchedArena
                for name in ("random", "greedy"):
                    arena = BatchedArena(self.nnet.predict_batch, int(self.args.numMCTSSims), self.args.cpuct, opponent=name,
                                         seed=int(np.random.randint(0, 2 ** 31 - 1)), max_turns=self.game._max_turns)
                    p1wins, p2wins, draws = arena.playGames(int(self.args.arenaCompare))
                    log.info(f"NEW/{name.capitalize()}Player WINS : {p1wins} / {p2wins} ; DRAWS : {draws}")
                continue
            player1 = MCTSPlayer(MCTS(self.nnet, self.args))
            for player2 in (RandomPlayer(), GreedyPlayer()):
                arena = Arena(player1, player2, self.game.restarted())
                p1wins, p2wins, draws = arena.playGames(self.args.arenaCompare)
                log.info(f"NEW/{player2.__class__.__name__} WINS : {p1wins} / {p2wins} ; DRAWS : {draws}")

    # ---- example files (Coach.py:167-193) -----------------------------------------------------------------
    def getCheckpointFile(self, iteration):
        return "checkpoint_" + str(iteration) + ".pth.tar"

    def saveTrainExamples(self, iteration):
        folder = self.args.checkpoint
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, self.getCheckpointFile(iteration) + ".examples"), "wb+") as f:
            Pickler(f).dump(self.trainExamplesHistory)

    def loadTrainExamples(self):
        model_file = os.path.join(self.args.load_folder_file[0], self.args.load_folder_file[1])
        examples_file = model_file + ".examples"
        if not os.path.isfile(examples_file):
            log.warning(f'File "{examples_file}" with trainExamples not found!')
            if input("Continue? [y|n]") != "y":
                sys.exit()
        else:
            with open(examples_file, "rb") as f:
                self.trainExamplesHistory = Unpickler(f).load()
            self.skipFirstSelfPlay = True

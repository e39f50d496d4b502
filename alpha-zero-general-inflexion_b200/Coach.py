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
    @staticmethod
    def _dist():
        """(rank, world) of the torch.distributed job, (0, 1) outside one."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _job_seed(self) -> int:
        """A fresh seed from numpy's global stream, the same on every rank of a torch.distributed job (rank 0's)."""
        seed = int(np.random.randint(0, 2 ** 31 - 1))
        rank, world = self._dist()
        if world > 1:
            import torch
            import torch.distributed as dist
            t = torch.tensor([seed], dtype=torch.int64, device=getattr(self.nnet, "device", "cpu"))
            dist.broadcast(t, 0)
            seed = int(t[0])
        return seed

    def _evaluator(self):
        """(evaluate, fused_head): on a GPU the net hands its raw binary16 head output to the tree kernel (softmax and
        tanh fused there); elsewhere the ordinary batched forward."""
        fused = hasattr(self.nnet, "predict_head") and getattr(getattr(self.nnet, "device", None), "type", "cpu") == "cuda"
        return (self.nnet.predict_head if fused else self.nnet.predict_batch), fused

    def selfPlayBatch(self, num_games: int, seed: int = 0, game_id_base: int = 0, gather: bool = False,
                      compact: bool = False):
        """``num_games`` episodes at once on the device -> list of (planes, pi list, v) in the reference's order.

        Under torch.distributed the games shard by rank in contiguous blocks (game i plays the same game whichever
        rank owns it; no collective during self-play) and every rank returns the examples of ITS games -- they stay
        rank-local and ``NNetWrapper.train`` all-reduces the gradients.  ``gather``: every rank
        returns the examples of ALL games instead (one all-gather of the un-expanded records, the "final example
        gather"; the 36-fold expansion happens after it).  ``compact``: return a ``selfplay.ExampleSet`` (same
        examples, expanded on demand per training batch) instead of the materialised list -- 1.5 KB per move instead
        of 19 MB per game, which is what lets an iteration hold thousands of games."""
        from .selfplay import ExampleSet, SelfPlayEngine, build_examples
        rank, world = self._dist()
        per = -(-int(num_games) // world)
        if world > 1 and per * world != num_games:
            raise ValueError(f"numEps={num_games} must be a multiple of the world size {world}")
        evaluate, fused = self._evaluator()
        eng = SelfPlayEngine(per, int(self.args.numMCTSSims), evaluate, cpuct=self.args.cpuct,
                             temp_threshold=int(self.args.tempThreshold), seed=seed, max_turns=self.game._max_turns,
                             game_id_base=game_id_base + rank * per, fused_head=fused,
                             nodes_per_game=dict(self.args).get("nodesPerGame", "auto"),
                             leaves_per_wave=int(dict(self.args).get("leavesPerWave", 1)))
        # the ranks stop together, as soon as every game everywhere is over (all-gather shapes agree; no rank keeps
        # searching finished games)
        eng.play_episode(collective=world > 1)
        eng.mcts.raise_on_faults()
        rec = eng.gather_records() if gather else eng.records()
        if compact:
            return ExampleSet(eng.be, rec)
        planes, pi, v = build_examples(eng.be, rec)
        return [(planes[i], pi[i].tolist(), float(v[i])) for i in range(len(v))]

    def learn(self):
        pit_interval = int(dict(self.args).get("pitInterval", 5))       # the reference hard-codes 5 (Coach.py:101,155)
        for i in range(1, self.args.numIters + 1):
            log.info(f"Starting Iter #{i} ...")
            rank, world = self._dist()
            opts = dict(self.args)
            compact = bool(opts.get("compactExamples", False))     # keep un-expanded records (selfplay.ExampleSet)
            if not self.skipFirstSelfPlay or i > 1:
                new = self.selfPlayBatch(int(self.args.numEps), seed=self._job_seed(),
                                         gather=bool(opts.get("gatherExamples", False)), compact=compact)
                if compact:
                    self.trainExamplesHistory.append(new)
                else:
                    examples = deque([], maxlen=self.args.maxlenOfQueue)
                    examples += new
                    self.trainExamplesHistory.append(examples)
            if len(self.trainExamplesHistory) > self.args.numItersForTrainExamplesHistory:
                log.warning(f"Removing the oldest entry in trainExamples. len(trainExamplesHistory) = "
                            f"{len(self.trainExamplesHistory)}")
                self.trainExamplesHistory.pop(0)
            if compact:                                            # (record files are not in the reference's format:
                from .selfplay import ExampleUnion                  # compact mode trades the .examples file for scale)
                self.nnet.train(ExampleUnion(self.trainExamplesHistory))
            else:
                self.saveTrainExamples(i - 1)
                train_examples = []
                for e in self.trainExamplesHistory:
                    train_examples.extend(e)
                shuffle(train_examples)
                self.nnet.train(train_examples)
            if rank == 0:
                self.nnet.save_checkpoint(folder=self.args.checkpoint, filename="temp.pth.tar")
            if i % pit_interval != 0:
                continue
            log.info("PITTING AGAINST BASELINES")
            if hasattr(self.nnet, "predict_batch"):
                # all arenaCompare games of a pairing at once on the device (same players, same colour split)
                # (sharded over the ranks of a torch.distributed job; only the W/L/D counters are all-reduced)
                from .arena_batched import BatchedArena
                evaluate, fused = self._evaluator()
                for name in ("random", "greedy"):
                    arena = BatchedArena(evaluate, int(self.args.numMCTSSims), self.args.cpuct, opponent=name,
                                         seed=self._job_seed(), max_turns=self.game._max_turns, fused_head=fused,
                                         leaves_per_wave=int(opts.get("leavesPerWave", 1)), capture=True)
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
        rank, world = self._dist()                     # rank-local shards under torch.distributed
        suffix = f".rank{rank}" if world > 1 else ""
        with open(os.path.join(folder, self.getCheckpointFile(iteration) + ".examples" + suffix), "wb+") as f:
            Pickler(f).dump(self.trainExamplesHistory)

    def loadTrainExamples(self):
        model_file = os.path.join(self.args.load_folder_file[0], self.args.load_folder_file[1])
        examples_file = model_file + ".examples"
        rank, world = self._dist()
        if world > 1 and os.path.isfile(examples_file + f".rank{rank}"):      # this rank's shard of a distributed run
            examples_file += f".rank{rank}"
        if not os.path.isfile(examples_file):
            log.warning(f'File "{examples_file}" with trainExamples not found!')
            if input("Continue? [y|n]") != "y":
                sys.exit()
        else:
            with open(examples_file, "rb") as f:
                self.trainExamplesHistory = Unpickler(f).load()
            self.skipFirstSelfPlay = True

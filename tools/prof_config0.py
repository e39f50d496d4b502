"""Dev tool (GPU box): cProfile of one Coach.executeEpisode episode through the drop-in classes (BASELINE configs[0])."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import inflexion_b200  # noqa: F401
from inflexion_b200.Coach import Coach
from inflexion_b200.MCTS import MCTS
from inflexion_b200.inflexion.InflexionGame import InflexionGame
from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
from inflexion_b200.utils import dotdict
np.random.seed(1); torch.manual_seed(0); torch.backends.cudnn.benchmark = True
game = InflexionGame(7, max_turns=343, max_power=6)
nnet = NNetWrapper(game)
args = dotdict({"numMCTSSims": 25, "cpuct": 1, "tempThreshold": 30})
coach = Coach(game, nnet, args)
coach.executeEpisode((game.restarted(), MCTS(nnet, args)))
pr = cProfile.Profile(); pr.enable()
ex = coach.executeEpisode((game.restarted(), MCTS(nnet, args)))
pr.disable()
print(len(ex) // 36, "moves")
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(30)
st.sort_stats("cumulative").print_stats(35)

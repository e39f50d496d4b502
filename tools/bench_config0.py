"""BASELINE configs[0] side by side: ONE self-play episode through Coach.executeEpisode (numMCTSSims = 25, tempThreshold = 30,
cpuct = 1, default CNN, 7x7 board, max_turns = 343 -- the reference's main.py settings).

  reference leg: the UNMODIFIED reference from oracle/_ref (its Coach, MCTS, InflexionGame, NNetWrapper) on the host CPU, in a
                 subprocess with CUDA hidden and torch's default thread count (the way `python main.py` runs it);
  ours leg:      the drop-in classes of this repo on cuda:0 -- the same call, Coach.executeEpisode((game, MCTS(nnet, args))).

Prints one JSON line per leg: seconds per episode, moves, examples, simulations/s.  The nets are random-init (no checkpoint
travels), so the two legs play different games of similar length; the figure of merit is seconds per move / simulations/s."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMS, TEMP_THRESHOLD, MAX_TURNS = 25, 30, int(os.environ.get("CFG0_MAX_TURNS", 343))


def reference_leg():
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import numpy as np
    import torch
    from Coach import Coach
    from MCTS import MCTS
    from inflexion.InflexionGame import InflexionGame
    from inflexion.pytorch.NNet import NNetWrapper
    from utils import dotdict
    np.random.seed(1)
    torch.manual_seed(0)
    game = InflexionGame(7, max_turns=MAX_TURNS, max_power=6)
    nnet = NNetWrapper(game)
    args = dotdict({"numMCTSSims": SIMS, "cpuct": 1, "tempThreshold": TEMP_THRESHOLD})
    coach = Coach(game, nnet, args)
    t0 = time.time()
    examples = coach.executeEpisode((game.restarted(), MCTS(nnet, args)))
    dt = time.time() - t0
    moves = len(examples) // 36
    print(json.dumps({"leg": "reference (unmodified, oracle/_ref, CPU)", "episodes": 1, "seconds_per_episode": dt, "moves": moves,
                      "examples": len(examples), "sims_per_s": moves * SIMS / dt, "ms_per_move": 1e3 * dt / moves,
                      "torch_threads": torch.get_num_threads(), "host_cores": os.cpu_count()}), flush=True)


def ours_leg(episodes=3):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import inflexion_b200  # noqa: F401
    from inflexion_b200.Coach import Coach
    from inflexion_b200.MCTS import MCTS
    from inflexion_b200.inflexion.InflexionGame import InflexionGame
    from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
    from inflexion_b200.utils import dotdict
    np.random.seed(1)
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    game = InflexionGame(7, max_turns=MAX_TURNS, max_power=6)
    nnet = NNetWrapper(game)
    for capture in (True, False):
        args = dotdict({"numMCTSSims": SIMS, "cpuct": 1, "tempThreshold": TEMP_THRESHOLD, "captureSearch": capture})
        coach = Coach(game, nnet, args)
        coach.executeEpisode((game.restarted(), MCTS(nnet, args)))            # warm-up episode (cuDNN heuristics, allocator)
        torch.cuda.synchronize()
        t0, moves, n_ex = time.time(), 0, 0
        for _ in range(episodes):
            ex = coach.executeEpisode((game.restarted(), MCTS(nnet, args)))
            moves, n_ex = moves + len(ex) // 36, n_ex + len(ex)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(json.dumps({"leg": "ours (drop-in classes, cuda:0, " + ("captured (net, wave) step" if capture else "eager loop") + ")",
                          "episodes": episodes, "seconds_per_episode": dt / episodes, "moves": moves / episodes,
                          "examples": n_ex / episodes, "sims_per_s": moves * SIMS / dt, "ms_per_move": 1e3 * dt / moves,
                          "device": torch.cuda.get_device_name(0)}), flush=True)


if __name__ == "__main__":
    leg = sys.argv[1] if len(sys.argv) > 1 else "both"
    if leg == "reference":
        reference_leg()
    elif leg == "ours":
        ours_leg()
    else:                                                    # separate processes: the two trees share module names
        for which in ("reference", "ours"):
            subprocess.run([sys.executable, os.path.abspath(__file__), which], check=False)

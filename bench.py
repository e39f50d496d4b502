#!/usr/bin/env python
"""bench.py -- self-play hot path throughput (MCTS simulations/sec; games/sec derived).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels + PyTorch net)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

One "step" = one self-play move for every concurrent game: numMCTSSims simulations per game
(tree-kernel waves interleaved with batched net forwards), the action choice and the root advance.
Workload (BASELINE.json configs[2]): 4096 concurrent games x 100 sims/move per GPU, default net
(4x conv 512 + fc, random init), synthetic = games played from the start position with the net's own priors.
Under torchrun (N>1) every rank runs its own shard of games (weak scaling, no data-path collective).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MOVES_PER_GAME = 344          # max_turns=343 -> at most 344 moves (games can end far earlier: a stack captured at ply 3 ends one);
                              # used only for the CPU arms' derived games/s, the GPU arm counts finished episodes
MEAN_EPISODE_MOVES = 301         # random-init default net, 4096 whole episodes (profiles/r1_soak_full_episode.txt)
TREE_BYTES_PER_VISIT = 4204   # P,N,Q rows 343*12 + 32 B header + 32 B hash probe + 24 B backup RMW
TREE_BYTES_PER_EXPANSION = 3596  # policy 1372 + v 4 + P row 1372 + header 32 + hash insert 32 + NN input 784
NET_FLOP_PER_EVAL = 404_287_488


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=4096, help="concurrent games per GPU")
    ap.add_argument("--sims", type=int, default=100, help="numMCTSSims per move")
    ap.add_argument("--net", default="default", choices=["default", "stub"],
                    help="default = the reference CNN (bf16 inference); stub = deterministic hash net (tree+rules only)")
    ap.add_argument("--workload", default="selfplay", choices=["selfplay", "rules"],
                    help="selfplay = the headline (BASELINE configs[2]); rules = the rules kernel alone (configs[1] shape)")
    ap.add_argument("--states", type=int, default=4 * 1024 * 1024, help="states per launch for --workload rules")
    ap.add_argument("--nodes", type=int, default=None, help="node pool slots per game (default: mcts_batched.default_pool_size)")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=1, choices=[1, 2],
                    help="2 = search the games as two halves on two streams inside the captured move (tree wave of one half "
                         "under the net forward of the other)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra-rooflines", action="store_true",
                    help="skip the stub-net tree roofline and the rules-kernel roofline appended to the default N=1 line")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU side.  Two implementations can play the same self-play loop on the host cores:
#   "reference": the UNMODIFIED reference (MCTS.py + inflexion/InflexionGame.py + inflexion/pytorch/NNet.py), staged
#                verbatim under oracle/_ref/ by oracle/make_ref.py and hash-checked before use;
#   "port":      the numpy oracle (oracle/inflexion_mcts.py), used when oracle/_ref is absent and as a second figure.
# One process per host core, one torch thread each, every process plays its own episodes and starts a new one when
# a game ends (continuous self-play, Coach.py:108-112 -- what the GPU arm's auto_restart does).
# ------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def reference_available() -> bool:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import make_ref
        return make_ref.verify(REF_DIR)
    except Exception:
        return False
    finally:
        sys.path.pop(0)


def _play_loop(play_one_move, sims, warm_moves, timed_moves, budget_s):
    """Runs `warm_moves` untimed moves, then EITHER exactly `timed_moves` moves OR moves until `budget_s` seconds have
    passed.  Returns (moves, sims, seconds, episodes_finished) of the timed part.  play_one_move() -> True when the
    game it just advanced is over (it must then start a new episode itself on the next call)."""
    for _ in range(warm_moves):
        play_one_move()
    moves, finished = 0, 0
    t0 = time.perf_counter()
    while True:
        finished += int(bool(play_one_move()))
        moves += 1
        dt = time.perf_counter() - t0
        if (timed_moves is not None and moves >= timed_moves) or (timed_moves is None and dt >= budget_s):
            return moves, moves * sims, dt, finished


def _stub_predict():
    import numpy as np
    import torch
    import inflexion_b200  # noqa: F401
    from inflexion_b200.stubnet import stub_policy_value

    def predict(planes):
        pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None])
        return pi.numpy()[0], v.numpy()[0]
    return predict


def _ref_worker(job):
    """One process of the reference arm: the reference's own classes, driven as Coach.executeEpisode drives them
    (Coach.py:66-82: getActionProb, the 36 symmetries of the example, np.random.choice, to_next_state)."""
    wid, sims, net_kind, warm_moves, timed_moves, budget_s, seed, max_turns = job
    os.environ["CUDA_VISIBLE_DEVICES"] = ""            # NNet.py:20 'cuda': torch.cuda.is_available() -> the CPU path
    sys.path.insert(0, REF_DIR)
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from MCTS import MCTS
    from flags import GameOutcome
    from inflexion.InflexionGame import InflexionGame
    from inflexion.pytorch.NNet import NNetWrapper
    from utils import dotdict
    np.random.seed(seed + wid)
    torch.manual_seed(0)
    game0 = InflexionGame(7, max_turns=max_turns, max_power=6)
    nnet = NNetWrapper(game0)
    if net_kind == "stub":                               # the deterministic evaluator of the parity contract
        nnet.predict = _stub_predict()
    args = dotdict({"numMCTSSims": sims, "cpuct": 1, "tempThreshold": 30})
    st = {"game": None, "mcts": None, "step": 0}

    def play_one_move():
        if st["game"] is None:                           # Coach.py:109-110
            st["game"], st["mcts"], st["step"] = game0.restarted(), MCTS(nnet, args), 0
        game, mcts = st["game"], st["mcts"]
        st["step"] += 1
        pi = mcts.getActionProb(game, temp=int(st["step"] < args.tempThreshold))
        game.symmetries(pi.reshape(game.policy_shape))   # Coach.py:77-78 (the example's 36 forms)
        game.symmetries(game.to_planes())
        game = game.to_next_state(np.random.choice(len(pi), p=pi))
        over = game.outcome != GameOutcome.ONGOING
        st["game"] = None if over else game
        return over
    return _play_loop(play_one_move, sims, warm_moves, timed_moves, budget_s)


def _port_worker(job):
    """One process of the port arm: the numpy oracle's restatement of the same loop."""
    wid, sims, net_kind, warm_moves, timed_moves, budget_s, seed, max_turns = job
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import numpy as np
    import torch
    torch.set_num_threads(1)
    import inflexion_b200  # noqa: F401
    from oracle.inflexion_mcts import PhiloxDraws, self_play_moves
    if net_kind == "default":
        from inflexion_b200.inflexion.pytorch.InflexionNNet import InflexionNNet
        from inflexion_b200.utils import dotdict

        class _Shape:
            board_shape, max_actions = (7, 7), 343
        torch.manual_seed(0)
        net = InflexionNNet(_Shape(), dotdict({"num_channels": 512, "dropout": 0.3})).eval()

        def predict(planes):                      # NNetWrapper.predict on CPU (NNet.py:78-94)
            with torch.no_grad():
                lp, v = net(torch.from_numpy(planes.astype(np.float32)).view(1, 4, 7, 7))
            return torch.exp(lp).numpy()[0], v.numpy()[0]
    else:
        predict = _stub_predict()
    st = {"it": None, "episodes": 0}

    def play_one_move():
        if st["it"] is None:                             # a new episode on a fresh RNG stream
            st["it"] = self_play_moves(predict, sims, 1, PhiloxDraws(seed, game=wid + 100003 * st["episodes"]),
                                       temp_threshold=30, max_turns=max_turns)
            st["episodes"] += 1
        rec = next(st["it"])
        over = rec["game"].outcome != 0
        if over:
            st["it"] = None
        return over
    return _play_loop(play_one_move, sims, warm_moves, timed_moves, budget_s)


def cpu_run(kind: str, sims: int, net_kind: str, warm_moves: int, timed_moves, budget_s, procs: int,
            max_turns: int = 343, seed: int = 1234) -> dict:
    """`procs` worker processes play concurrently; -> aggregate of their timed parts.  value = sum over workers of
    (sims / seconds): every worker is timed on its own clock while all of them load the cores together."""
    import multiprocessing as mp
    worker = _ref_worker if kind == "reference" else _port_worker
    jobs = [(w, sims, net_kind, warm_moves, timed_moves, budget_s, seed, max_turns) for w in range(procs)]
    if procs == 1:
        res = [worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(worker, jobs)
    moves = sum(r[0] for r in res)
    total_sims = sum(r[1] for r in res)
    secs = [r[2] for r in res]
    return {"value": sum(r[1] / r[2] for r in res), "moves": moves, "sims": total_sims, "seconds_max": max(secs),
            "seconds_mean": sum(secs) / len(secs), "episodes_finished": sum(r[3] for r in res),
            "moves_per_worker": moves / procs}


def _cpu_kind() -> str:
    return "reference" if reference_available() else "port"


def _cpu_sample_text(kind, procs, r, sims, net_kind):
    what = ("the UNMODIFIED reference (MCTS.py + InflexionGame.py + NNetWrapper on CPU, staged verbatim in oracle/_ref)"
            if kind == "reference" else "the numpy oracle port of the reference MCTS + rules")
    return (f"{procs} processes (one per host core, 1 torch thread each) x {r['moves_per_worker']:.1f} self-play moves x "
            f"{sims} sims of {what}, {net_kind} net, continuous episodes; {r['seconds_max']:.1f} s timed per worker after "
            f"warm-up; value = sum of per-worker sims/s")


def cpu_baseline(sims: int, net_kind: str, budget_s: float) -> dict:
    """The CPU figure reported beside the GPU line: a bounded sample (about `budget_s` seconds per worker after one
    warm-up move).  Never raises: a failure of this leg must not cost the GPU line."""
    procs = os.cpu_count() or 1
    kind = _cpu_kind()
    try:
        r = cpu_run(kind, sims, net_kind, 1, None, budget_s, procs)
        out = {"value": r["value"], "unit": "sims/s", "cores": procs, "kind": kind,
               "sample": _cpu_sample_text(kind, procs, r, sims, net_kind)}
    except Exception as e:                                  # noqa: BLE001
        return {"value": None, "unit": "sims/s", "cores": procs, "kind": kind, "error": f"{type(e).__name__}: {e}"[:300]}
    if kind == "reference":                                 # the oracle port as a second, labelled figure
        try:
            p = cpu_run("port", sims, net_kind, 1, None, max(3.0, budget_s / 3), procs)
            out["port"] = {"value": p["value"], "unit": "sims/s", "kind": "port",
                           "sample": _cpu_sample_text("port", procs, p, sims, net_kind)}
        except Exception as e:                              # noqa: BLE001
            out["port"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, same metric/config.
    A step = one self-play move (a.sims simulations) in each of the worker games."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:                                           # under torchrun only rank 0 runs the CPU arm
        return
    procs = os.cpu_count() or 1
    kind = _cpu_kind()
    r = cpu_run(kind, a.sims, a.net, a.warmup, a.steps, None, procs)
    value = r["value"]
    line = {
        "impl": "reference", "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * r["seconds_mean"] / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "games_per_s": value / (a.sims * MEAN_EPISODE_MOVES),
        "config": workload_config(a, a.games, note=f"CPU arm: bounded sample of this workload, {procs} games at a time (one per "
                                                    f"host core) x {a.sims} sims per step"),
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": procs, "kind": kind,
                         "sample": _cpu_sample_text(kind, procs, r, a.sims, a.net)},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


LARGE_LAUNCH_GAMES, LARGE_LAUNCH_NODES = 16384, 640


def ncu_traffic(kernel: str, workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    of this same command (profiles/traffic.json, written by tools/summarize_profiles.py); None if not captured."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel][workload]["dram_bytes_per_launch"]
    except Exception:
        return None


def workload_config(a, per_gpu_games, note=None):
    cfg = {"workload": f"batched Inflexion self-play, {per_gpu_games} concurrent games x {a.sims} sims/move per GPU, "
                       f"{a.net} net, 7x7 board, max_turns=343, cpuct=1, tempThreshold=30",
           "games_per_gpu": per_gpu_games, "sims_per_move": a.sims, "net": a.net,
           "arithmetic": "tree: float32 with numpy's exact operation order (float64 where the reference's Python floats are)"
                         + ("; net: fp16 inputs/weights, fp32 accumulate" if a.net == "default" else "; net: integer-hash stub"),
           "l2_policy": "working set (node pools) far larger than L2; no flush needed",
           "parallelism": f"games sharded over {a.gpus} GPU(s), no collective"}
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import inflexion_b200  # noqa: F401
    from inflexion_b200.backend import Backend
    from inflexion_b200.selfplay import SelfPlayEngine
    from inflexion_b200.stubnet import stub_kernel_evaluator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")         # (keeps NCCL's version banner off stdout: one JSON line only)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = Backend.cuda(local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured" if peaks else "fallback"

    if a.net == "default":
        from inflexion_b200.inflexion.InflexionGame import InflexionGame
        from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
        torch.manual_seed(0)
        torch.backends.cudnn.benchmark = True
        nnet = NNetWrapper(InflexionGame(7, max_turns=343, max_power=6), device=be.device)
        evaluate, fused = nnet.predict_head, True          # last GEMM -> tree kernel (softmax/tanh fused there)
    else:
        evaluate, fused = None, False                      # installed below: the stub is a kernel of the library

    eng = SelfPlayEngine(a.games, a.sims, evaluate, cpuct=1.0, temp_threshold=30, seed=2023, max_turns=343,
                         game_id_base=rank * a.games, record=True, backend=be, fused_head=fused, nodes_per_game=a.nodes,
                         auto_restart=True, restart_stride=world * a.games, streams=a.streams)
    if a.net == "stub":                                    # deterministic stub evaluator, one kernel per wave (csrc/ifx_stub.cu)
        eng.evaluate = stub_kernel_evaluator(eng.mcts)
    eng.extra_launches_per_search = a.sims                 # per net evaluation one kernel of ours: the stub, or conv1's im2col

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the CUDA graph) ------------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        eng.play_move()
    eng.mcts.raise_on_faults()
    # ---- timed region: EXACTLY K steps, device-resident, one CUDA-graph launch per move -------------------------
    l0 = eng.my_launches
    c0 = eng.mcts.counters()
    r0 = int(eng.restarts.item())
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        t_start.record()
        for _ in range(a.steps):
            eng.play_move()
        t_end.record()
        barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    c1 = eng.mcts.counters()
    eng.mcts.raise_on_faults()
    launches = eng.my_launches - l0
    if world > 1:
        t = torch.tensor([elapsed_ms], device=be.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t[0])
    sims_rank = c1["sims"] - c0["sims"]
    if world > 1:                                      # the units all ranks processed (each rank counts its own on the device)
        t = torch.tensor([float(sims_rank)], device=be.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        sims_total = float(t[0])
    else:
        sims_total = float(sims_rank)
    value = sims_total / (elapsed_ms * 1e-3)

    restarts_timed = int(eng.restarts.item()) - r0
    if world > 1:
        t = torch.tensor([float(restarts_timed)], device=be.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        restarts_total = float(t[0])
    else:
        restarts_total = float(restarts_timed)

    # ---- instrumented pass: the same K steps again with every tree wave launched outside the graph and bracketed by
    # CUDA events recorded inside the library call on the launching stream (roofline of OUR dominant kernel) -------
    tree = tree_roofline(eng, a.steps, hbm_peak, peak_src, ncu_traffic("tree_wave_kernel", a.net) if (a.games, a.sims) == (4096, 100) else None)
    net_ms = tree.pop("_instr_ms") - tree.pop("_tree_ms")
    net_evals = tree.pop("_expansions")                    # one net row is consumed per expansion (counted by the kernel)
    net_tflops = net_evals * NET_FLOP_PER_EVAL / (net_ms * 1e-3) / 1e12 if a.net == "default" and net_ms > 0 else None

    # ---- end to end through the public host-buffer API: K steps, pinned host roots in, host results out, every step ----
    e2e = None
    if not a.no_e2e:
        roots_h = torch.empty((a.games, 64), dtype=torch.uint8).pin_memory()
        next_h = torch.empty((a.games, 64), dtype=torch.uint8).pin_memory()
        counts_h = torch.empty((a.games, 343), dtype=torch.int32).pin_memory()
        acts_h = torch.empty((a.games,), dtype=torch.int32).pin_memory()
        roots_h.copy_(eng.mcts.roots())
        torch.cuda.synchronize()
        restart_h = torch.zeros((a.games,), dtype=torch.int32).pin_memory()
        start_state = torch.from_numpy(np.ascontiguousarray(
            __import__("inflexion_b200").state.pack_states(np.zeros((1, 49), dtype=np.int64), 1, 0, 0, 343)).view(np.uint8).reshape(64))
        next_id = [(world + rank) * a.games * 1000]

        def hand_over():
            """host side of the loop: next positions become the roots; finished games start a new episode"""
            roots_h.copy_(next_h)
            restart_h.copy_(roots_h[:, 45] != 0)
            over = torch.nonzero(restart_h).flatten()
            for g in over.tolist():
                roots_h[g] = start_state
                roots_h[g, 48:52] = torch.from_numpy(np.array([next_id[0]], dtype=np.uint32).view(np.uint8))
                next_id[0] += 1
        eng.step_host(roots_h, counts_h, acts_h, next_h, restart_h)   # warm
        hand_over()
        barrier()
        ce0 = eng.mcts.counters()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.step_host(roots_h, counts_h, acts_h, next_h, restart_h)
            hand_over()
        barrier()
        dt = time.perf_counter() - t0
        ce1 = eng.mcts.counters()
        if world > 1:
            t = torch.tensor([dt], device=be.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        e2e_sims = (ce1["sims"] - ce0["sims"]) * world
        e2e = {"value": e2e_sims / dt, "unit": "sims/s", "h2d_bytes_per_step": a.games * (64 + 4),
               "d2h_bytes_per_step": a.games * (343 * 4 + 4 + 64), "steps": a.steps,
               "api": "SelfPlayEngine.step_host(pinned host roots) -> host visit counts, actions, next roots"}

    # ---- the other two kernels of the path, measured in the same run (rank 0, N=1: they do not shard differently) ----
    extra = {}
    if world == 1 and a.net == "default" and not a.no_extra_rooflines:
        del eng
        torch.cuda.empty_cache()

        def stub_tree(games, nodes, traffic, workload):
            """tree_roofline of a fresh engine with the deterministic stub evaluator (one kernel of the library)."""
            try:
                stub = SelfPlayEngine(games, a.sims, None, cpuct=1.0, temp_threshold=30, seed=2023, max_turns=343, record=False,
                                      backend=be, fused_head=False, nodes_per_game=nodes, auto_restart=True)
                stub.evaluate = stub_kernel_evaluator(stub.mcts)
                for _ in range(3):
                    stub.play_move()
                r = tree_roofline(stub, max(3, min(a.steps, 8)), hbm_peak, peak_src, traffic)
                stub.mcts.raise_on_faults()
                for k in ("_instr_ms", "_tree_ms", "_expansions"):
                    r.pop(k)
                r["workload"], r["games_per_launch"] = workload, games
                del stub
                return r
            except Exception as e:                          # noqa: BLE001
                return {"error": f"{type(e).__name__}: {e}"[:300]}
            finally:
                torch.cuda.empty_cache()
        extra["roofline_tree_stub"] = stub_tree(
            a.games, a.nodes, ncu_traffic("tree_wave_kernel", "stub") if (a.games, a.sims) == (4096, 100) else None,
            "same games x sims with the deterministic stub evaluator (tree + rules only, deeper trees)")
        # One warp per game and 4 096 warps fit on the 148 SMs at once, so a 4 096-game launch lasts as long as its DEEPEST
        # descent (profiles/r2_trace_stub.txt: mean warp lifetime 10.7 us, longest 16.9 us).  With more games per launch than
        # resident warps the CTAs that finish early make room for the rest and the tail is amortised: the same kernel,
        # 4x the games (node pool of 640 slots per game so that it fits beside nothing else: 74 GB).
        if (a.games, a.sims) == (4096, 100):
            extra["roofline_tree_stub_16k"] = stub_tree(
                LARGE_LAUNCH_GAMES, LARGE_LAUNCH_NODES, ncu_traffic("tree_wave_kernel", "stub16k"),
                f"{LARGE_LAUNCH_GAMES} games per launch (4x the headline's; more warps than the SMs hold at once), stub evaluator")
        try:
            extra["roofline_rules"] = rules_roofline(be, a.states, max(a.steps, 10), hbm_peak, peak_src)
        except Exception as e:                              # noqa: BLE001
            extra["roofline_rules"] = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32-tree/f16-net" if a.net == "default" else "f32",
            "data": "synthetic",
            # episodes that ended inside the timed region (counted on the device by the restart kernel) per second, and the
            # steady-state figure moves/s / mean episode length
            "games_per_s": value / (a.sims * MEAN_EPISODE_MOVES),
            "games_per_s_note": f"moves/s / {MEAN_EPISODE_MOVES} (mean episode length of this net, measured over whole episodes by "
                                "tools/soak_episode.py: profiles/r2_soak_full_episode.txt); episodes that ENDED inside the timed "
                                "region (counted on the device by the restart kernel) are in games_finished_in_timed_region",
            "games_finished_in_timed_region": restarts_total,
            "config": workload_config(a, a.games),
            "roofline": tree,
            "roofline_net": None if net_tflops is None else {
                "kernel": "policy-value net forward (cuDNN/cuBLAS fp16, library)", "bound": "tensor", "achieved": net_tflops,
                "peak": tensor_peak, "unit": "TFLOP/s", "frac": net_tflops / tensor_peak,
                "net_evals": net_evals,
                "note": "library kernels, reported for context only: (instrumented pass - tree waves) / evaluations consumed "
                        "(= expansions counted by the tree kernel); peak = measured sustained bf16 GEMM rate"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        }
        line.update(extra)
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a.sims, a.net, a.cpu_baseline_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def tree_roofline(eng, steps, hbm_peak, peak_src, traffic):
    """Instrumented pass over `steps` self-play moves of `eng`: every tree wave is launched outside the CUDA graph and
    bracketed by a CUDA event pair recorded inside ifx_mcts_wave on the launching stream.

    Two conventions, both reported: `achieved`/`frac` = ALGORITHMIC bytes (the reference's dense-row scan:
    4204 B per node visit + 3596 B per expansion, both counted by the kernel) / launch time; `achieved_dram`/`frac_dram` =
    the DRAM bytes ncu saw for the same launch (profiles/traffic.json) / the same launch time.  Where the ncu bytes are
    below the formula (compact rows, on-chip reuse) the survey makes the second the binding figure: `frac_binding`."""
    import torch
    eng.mcts.wave_timing(True)
    eng.play_move(True)                                    # warm the un-captured path
    eng.mcts.wave_time()
    c2 = eng.mcts.counters()
    torch.cuda.synchronize()
    i_start, i_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i_start.record()
    tree_ms, n_tree = 0.0, 0
    for k in range(steps):
        eng.play_move(True)
        if (k + 1) % 30 == 0:                              # drain the library's event ring (4096 pairs) before it fills
            ms, n = eng.mcts.wave_time()
            tree_ms, n_tree = tree_ms + ms, n_tree + n
    i_end.record()
    torch.cuda.synchronize()
    ms, n = eng.mcts.wave_time()
    tree_ms, n_tree = tree_ms + ms, n_tree + n
    eng.mcts.wave_timing(False)
    instr_ms = i_start.elapsed_time(i_end)
    c3 = eng.mcts.counters()
    visits, expansions = c3["node_visits"] - c2["node_visits"], c3["expansions"] - c2["expansions"]
    sims = max(c3["sims"] - c2["sims"], 1)
    tree_bytes = TREE_BYTES_PER_VISIT * visits + TREE_BYTES_PER_EXPANSION * expansions
    gbs = tree_bytes / (tree_ms * 1e-3) / 1e9 if tree_ms > 0 else 0.0
    us = 1e3 * tree_ms / max(n_tree, 1)
    dram = None if not traffic or tree_ms <= 0 else traffic / (1e3 * us)
    alg_per_launch = tree_bytes / max(n_tree, 1)
    return {"kernel": "tree_wave_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
            "achieved_dram": dram, "frac_dram": None if dram is None else dram / hbm_peak,
            "frac_binding": (gbs / hbm_peak) if (dram is None or traffic >= alg_per_launch) else dram / hbm_peak,
            "convention": "frac = algorithmic bytes (dense-row scan of the reference: 4204 B/visit + 3596 B/expansion, counted by "
                          "the kernel) / event-timed launch; frac_dram = ncu DRAM bytes of the committed capture / the same time; "
                          "frac_binding = frac_dram where ncu bytes < algorithmic bytes (SURVEY.md 8d), else frac",
            "algorithmic_bytes_per_launch": alg_per_launch, "avg_launch_us": us,
            "launches": n_tree, "share_of_step": tree_ms / instr_ms,
            "measured_in": "instrumented pass (waves outside the CUDA graph, event pairs recorded inside ifx_mcts_wave)",
            "visits_per_sim": visits / sims, "expansions_per_sim": expansions / sims, "sims_per_launch": sims / max(n_tree, 1),
            "_instr_ms": instr_ms, "_tree_ms": tree_ms, "_expansions": expansions}


def rules_states(be, n):
    """n uniform-random reachable positions (BASELINE configs[1]): random play from the start for a random number of plies."""
    import numpy as np
    import torch
    from inflexion_b200 import rules
    from inflexion_b200.state import pack_states
    rng = np.random.default_rng(0)
    plies = torch.from_numpy(rng.integers(0, 344, size=n)).to(be.device)
    states = be.states_tensor(pack_states(np.zeros((n, 49), dtype=np.int64), 1, 0, 0, 343, tags=np.arange(n)))
    for t in range(343):
        acts = rules.random_actions(states, 1, t, backend=be)
        _, nxt, status = rules.mask_step(states, acts, want_masks=False, backend=be)
        states = torch.where(((plies > t) & (status == 0) & (nxt[:, 45] == 0))[:, None], nxt, states)
    return states, rules.random_actions(states, 2, 0, backend=be)


def rules_roofline(be, n, steps, hbm_peak, peak_src, clocks_index=None):
    """rules_mask_step_kernel alone: valid-move mask + step for n states per launch, one CUDA event pair per launch."""
    import torch
    states, acts = rules_states(be, n)
    masks, nxt, status = (be.empty((n, 7), torch.int64), be.empty((n, 64), torch.uint8), be.empty((n,), torch.int32))
    stream = be.stream()

    def launch():
        be.check(be.dll.ifx_rules_mask_step(states.data_ptr(), acts.data_ptr(), n, masks.data_ptr(), nxt.data_ptr(),
                                            status.data_ptr(), stream))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); launch(); e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    ms = [s.elapsed_time(e) for s, e in evs]
    avg = sum(ms) / len(ms)
    bytes_per_launch = 188 * n            # 64 B state in + 4 B action + 64 B state out + 56 B packed mask
    gbs = bytes_per_launch / (avg * 1e-3) / 1e9
    return {"kernel": "rules_mask_step_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": gbs / hbm_peak, "traffic": ncu_traffic("rules_mask_step_kernel", "rules"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": bytes_per_launch, "states_per_launch": n, "states_per_s": n / (avg * 1e-3),
            "avg_launch_us": avg * 1e3, "min_launch_us": min(ms) * 1e3, "launches": steps,
            "workload": f"{n} random reachable states per launch ({bytes_per_launch / 1e6:.0f} MB algorithmic, larger than L2); "
                        "188 B/state = 64 in + 4 action + 64 out + 56 packed mask"}


def run_rules(a):
    """Rules kernel alone: valid-move mask + step for `states` random reachable positions per launch."""
    import torch  # noqa: F401
    import inflexion_b200  # noqa: F401
    from inflexion_b200.backend import Backend
    be = Backend.cuda(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    with ClockSampler(0) as clocks:
        r = rules_roofline(be, a.states, a.steps, hbm_peak, "measured" if peaks else "fallback")
    print(json.dumps({
        "metric": "rules_states_per_sec", "value": r["states_per_s"], "unit": "states/s", "n_gpus": 1, "steps": a.steps,
        "warmup": 3, "ms_per_step": r["avg_launch_us"] * 1e-3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 bitboards", "data": "synthetic", "config": {"workload": "rules kernel: move mask + step, " + r["workload"]},
        "roofline": r, "gpu_launches": a.steps, "clocks": clocks.summary()}), flush=True)


if __name__ == "__main__":
    args = parse_args()
    if args.workload == "rules":
        run_rules(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

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

MOVES_PER_GAME = 344          # max_turns=343 -> 344 moves; random-init and stub nets always play the full length
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
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
# CPU side: the oracle port of the reference's MCTS + rules (+ the same CNN on CPU), all host cores
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process: plays self-play moves with the oracle MCTS for `moves` moves; returns (sims, seconds per move list)."""
    wid, sims, net_kind, moves, seed = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    import inflexion_b200  # noqa: F401
    from inflexion_b200.stubnet import stub_policy_value
    from oracle.inflexion_mcts import PhiloxDraws, self_play_moves
    if net_kind == "default":
        from inflexion_b200.inflexion.InflexionGame import InflexionGame  # noqa: F401  (shape source only)
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
        def predict(planes):
            pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None])
            return pi.numpy()[0], v.numpy()[0]
    times = []
    it = self_play_moves(predict, sims, 1, PhiloxDraws(seed, game=wid), temp_threshold=30, max_moves=moves)
    t0 = time.perf_counter()
    for _ in it:
        t1 = time.perf_counter()
        times.append(t1 - t0)
        t0 = t1
    return times


def cpu_run(sims: int, net_kind: str, moves: int, procs: int):
    """`procs` independent games, `moves` moves each, in parallel processes.  Returns per-move wall times of the
    slowest worker pattern: list over moves of max-over-workers time."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(w, sims, net_kind, moves, 1234) for w in range(procs)])
    per_move = [max(r[i] for r in res) for i in range(moves)]
    return per_move


def cpu_baseline(sims: int, net_kind: str, budget_s: float):
    procs = os.cpu_count() or 1
    # calibrate with one move, then size the sample to the budget
    t0 = time.perf_counter()
    first = cpu_run(sims, net_kind, 1, procs)
    spawn_and_move = time.perf_counter() - t0
    moves = max(1, min(6, int(budget_s / max(first[0], 1e-3))))
    per_move = cpu_run(sims, net_kind, moves, procs) if moves > 1 else first
    steady = per_move[1:] if len(per_move) > 1 else per_move
    total = sum(steady)
    value = procs * sims * len(steady) / total
    return {"value": value, "unit": "sims/s", "cores": procs, "kind": "port",
            "sample": f"{procs} processes x {len(steady)} self-play moves x {sims} sims, oracle MCTS+rules (numpy port of "
                      f"the reference) with the {net_kind} net on 1 torch thread each; first move excluded "
                      f"(calibration {spawn_and_move:.1f}s)",
            "games_per_s": value / (sims * MOVES_PER_GAME)}


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port) with all host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    moves = a.warmup + a.steps
    # a step = one self-play move (sims simulations) in each of `procs` worker games; bounded by design
    per_move = cpu_run(a.sims, a.net, moves, procs)
    timed = per_move[a.warmup:]
    total = sum(timed)
    value = procs * a.sims * len(timed) / total
    line = {
        "impl": "reference", "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "games_per_s": value / (a.sims * MOVES_PER_GAME),
        "config": workload_config(a, a.games, note=f"CPU arm: bounded sample of this workload, {procs} games (one per host core) "
                                                    f"x {a.sims} sims per step"),
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": procs, "kind": "port",
                         "sample": f"{procs} processes x {len(timed)} moves x {a.sims} sims (oracle port of the reference "
                                   f"MCTS+rules, {a.net} net on CPU, 1 torch thread per process)"},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
                         auto_restart=True, restart_stride=world * a.games)
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
    sims_total = sims_rank * world                     # every rank runs the same amount of work (weak scaling)
    value = sims_total / (elapsed_ms * 1e-3)

    # ---- instrumented pass: the same K steps again with every tree wave launched outside the graph and bracketed by
    # CUDA events recorded inside the library call on the launching stream (roofline of OUR dominant kernel) -------
    eng.mcts.wave_timing(True)
    eng.play_move(True)                                    # warm the un-captured path
    eng.mcts.wave_time()
    c2 = eng.mcts.counters()
    torch.cuda.synchronize()
    i_start, i_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i_start.record()
    tree_ms, n_tree = 0.0, 0
    for k in range(a.steps):
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
    sims_instr = c3["sims"] - c2["sims"]
    tree_bytes = TREE_BYTES_PER_VISIT * visits + TREE_BYTES_PER_EXPANSION * expansions
    tree_gbs = tree_bytes / (tree_ms * 1e-3) / 1e9 if tree_ms > 0 else 0.0
    net_ms = instr_ms - tree_ms
    net_evals = a.games * a.sims * a.steps
    net_tflops = net_evals * NET_FLOP_PER_EVAL / (net_ms * 1e-3) / 1e12 if a.net == "default" and net_ms > 0 else None
    sims_rank = max(sims_instr, 1)

    # ---- end to end through the public host-buffer API ----------------------------------------------------
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
        k2 = max(2, min(a.steps, 4))
        ce0 = eng.mcts.counters()
        t0 = time.perf_counter()
        for _ in range(k2):
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
               "d2h_bytes_per_step": a.games * (343 * 4 + 4 + 64), "steps": k2,
               "api": "SelfPlayEngine.step_host(pinned host roots) -> host visit counts, actions, next roots"}

    tree_traffic = ncu_traffic("tree_wave_kernel", a.net) if (a.games, a.sims) == (4096, 100) else None
    if rank == 0:
        line = {
            "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "games_per_s": value / (a.sims * MOVES_PER_GAME),
            "config": workload_config(a, a.games),
            "roofline": {"kernel": "tree_wave_kernel", "bound": "hbm", "achieved": tree_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": tree_gbs / hbm_peak, "traffic": tree_traffic, "peak_source": peak_src,
                         # the same launch time against the DRAM bytes ncu saw (compact rows + on-chip reuse move far
                         # fewer bytes than the reference's dense-row scan the algorithmic figure counts)
                         "achieved_dram": None if not tree_traffic or tree_ms <= 0 else tree_traffic / (1e6 * tree_ms / max(n_tree, 1)),
                         "frac_dram": None if not tree_traffic or tree_ms <= 0 else tree_traffic / (1e6 * tree_ms / max(n_tree, 1)) / hbm_peak,
                         "algorithmic_bytes_per_launch": tree_bytes / max(n_tree, 1), "avg_launch_us": 1e3 * tree_ms / max(n_tree, 1),
                         "launches": n_tree, "share_of_step": tree_ms / instr_ms,
                         "measured_in": "instrumented pass of the same K steps (waves outside the CUDA graph, event pairs recorded "
                                        "inside ifx_mcts_wave)",
                         "visits_per_sim": visits / max(sims_rank, 1), "expansions_per_sim": expansions / max(sims_rank, 1)},
            "roofline_net": None if net_tflops is None else {
                "kernel": "policy-value net forward (cuDNN/cuBLAS fp16, library)", "bound": "tensor", "achieved": net_tflops,
                "peak": tensor_peak, "unit": "TFLOP/s", "frac": net_tflops / tensor_peak, "share_of_step": net_ms / instr_ms,
                "note": "everything in the instrumented pass that is not the tree wave; peak = measured sustained bf16 GEMM rate (fp16 has the same nominal rate)"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(),
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a.sims, a.net, a.cpu_baseline_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_rules(a):
    """Rules kernel alone: valid-move mask + step for `states` random reachable positions per launch."""
    import numpy as np
    import torch
    import inflexion_b200  # noqa: F401
    from inflexion_b200 import rules
    from inflexion_b200.backend import Backend
    from inflexion_b200.state import pack_states
    be = Backend.cuda(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n = a.states
    rng = np.random.default_rng(0)
    plies = torch.from_numpy(rng.integers(0, 344, size=n)).to(be.device)
    states = be.states_tensor(pack_states(np.zeros((n, 49), dtype=np.int64), 1, 0, 0, 343, tags=np.arange(n)))
    for t in range(343):                                     # uniform-random reachable positions (config 2)
        acts = rules.random_actions(states, 1, t, backend=be)
        _, nxt, status = rules.mask_step(states, acts, want_masks=False, backend=be)
        states = torch.where(((plies > t) & (status == 0) & (nxt[:, 45] == 0))[:, None], nxt, states)
    acts = rules.random_actions(states, 2, 0, backend=be)
    masks, nxt, status = (be.empty((n, 7), torch.int64), be.empty((n, 64), torch.uint8), be.empty((n,), torch.int32))
    stream = be.stream()

    def launch():
        be.check(be.dll.ifx_rules_mask_step(states.data_ptr(), acts.data_ptr(), n, masks.data_ptr(), nxt.data_ptr(),
                                            status.data_ptr(), stream))
    for _ in range(max(a.warmup, 3)):
        launch()
    torch.cuda.synchronize()
    evs = []
    with ClockSampler(0) as clocks:
        for _ in range(a.steps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); launch(); e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
    ms = [s.elapsed_time(e) for s, e in evs]
    avg = sum(ms) / len(ms)
    bytes_per_launch = 188 * n            # 64 B state in + 4 B action + 64 B state out + 56 B packed mask
    gbs = bytes_per_launch / (avg * 1e-3) / 1e9
    print(json.dumps({
        "metric": "rules_states_per_sec", "value": n / (avg * 1e-3), "unit": "states/s", "n_gpus": 1, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": avg, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 bitboards", "data": "synthetic",
        "config": {"workload": f"rules kernel: move mask + step, {n} random reachable states per launch "
                               f"({bytes_per_launch / 1e6:.0f} MB per launch, larger than L2)"},
        "roofline": {"kernel": "rules_mask_step_kernel", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": gbs / hbm_peak, "traffic": ncu_traffic("rules_mask_step_kernel", "rules"),
                     "algorithmic_bytes_per_launch": bytes_per_launch,
                     "avg_launch_us": avg * 1e3, "min_launch_us": min(ms) * 1e3},
        "gpu_launches": a.steps, "clocks": clocks.summary()}), flush=True)


if __name__ == "__main__":
    args = parse_args()
    if args.workload == "rules":
        run_rules(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

"""bench.py's GPU arm, dry-run on the CPU: torch.cuda is stubbed out and the kernels' logic-check build stands in for
the library, with a tiny workload.  This checks the CONTROL FLOW and the JSON line bench.py prints (every key the bench
contract names, sane types), not any number -- the numbers come from the B200 runs under profiles/."""
import json
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 1.0


def test_gpu_arm_control_flow_and_json_line(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    import inflexion_b200.backend as backend_mod
    from hostsim import hostsim_backend
    be = hostsim_backend()
    monkeypatch.setattr(backend_mod.Backend, "cuda", staticmethod(lambda index=None: be))
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="ours", games=6, sims=5, net="stub", workload="selfplay",
                                 states=0, nodes=None, cpu_baseline_seconds=0.0, no_cpu_baseline=True, no_e2e=False,
                                 no_extra_rooflines=True, streams=1)
    bench.run_ours(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["metric"] == "mcts_simulations_per_sec" and line["unit"] == "sims/s" and line["n_gpus"] == 1
    assert line["steps"] == 2 and line["warmup"] == 3 and line["scaling"] == "weak" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["steps"] == 2                  # e2e is timed over all K steps
    assert line["e2e"]["h2d_bytes_per_step"] == 6 * 68 and line["e2e"]["d2h_bytes_per_step"] == 6 * (343 * 4 + 4 + 64)
    r = line["roofline"]
    for key in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic", "achieved_dram", "frac_dram",
                "algorithmic_bytes_per_launch", "avg_launch_us", "launches", "visits_per_sim", "expansions_per_sim"):
        assert key in r, key
    assert r["kernel"] == "tree_wave_kernel" and r["bound"] == "hbm"
    assert r["launches"] in (0, 2 * 6)                # 2 steps x (sims + 1) waves on a GPU; the CPU build records no events
    assert r["traffic"] is None                       # only the captured workload (4096 x 100) carries the ncu figure
    # 2 timed steps: (sims+1) waves + sims stub-evaluator kernels + record/sample/advance/restart kernels each
    assert line["gpu_launches"] >= 2 * (6 + 5)
    assert line["config"]["games_per_gpu"] == 6 and "workload" in line["config"]


# ---- the CPU legs (cpu_baseline / --impl reference) ------------------------------------------------------------------
def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_cpu_worker_survives_games_that_end_early(kind):
    """Round 1's bench crashed on hosts with > 16 cores: game id 16 (seed 1234) ends after 3 moves and the per-move
    indexing assumed full-length games.  Workers now start a new episode when one ends: with max_turns=2 EVERY game
    ends after 3 moves, and the worker must still deliver exactly the moves it was asked for."""
    bench = _bench()
    if kind == "reference" and not bench.reference_available():
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py needs /root/reference)")
    worker = bench._ref_worker if kind == "reference" else bench._port_worker
    old = os.environ.get("CUDA_VISIBLE_DEVICES")
    try:
        moves, sims, secs, finished = worker((16, 4, "stub", 2, 9, None, 1234, 2))
    finally:
        if old is None:
            os.environ.pop("CUDA_VISIBLE_DEVICES", None)
        else:
            os.environ["CUDA_VISIBLE_DEVICES"] = old
        if bench.REF_DIR in sys.path:
            sys.path.remove(bench.REF_DIR)
        for m in ("MCTS", "Game", "flags", "utils", "NeuralNet", "inflexion", "inflexion.InflexionGame",
                  "inflexion.pytorch", "inflexion.pytorch.NNet", "inflexion.pytorch.InflexionNNet", "inflexion.utils"):
            if kind == "reference":
                sys.modules.pop(m, None)
    assert moves == 9 and sims == 36 and secs > 0
    assert finished >= 2                              # 11 moves of 3-move games: at least two of the timed ones ended a game


def test_cpu_run_aggregates_and_baseline_never_raises(monkeypatch):
    bench = _bench()
    r = bench.cpu_run("port", 3, "stub", 1, 4, None, 1, max_turns=2)
    assert r["moves"] == 4 and r["sims"] == 12 and r["value"] > 0 and r["episodes_finished"] >= 1
    monkeypatch.setattr(bench, "cpu_run", lambda *a, **k: (_ for _ in ()).throw(IndexError("boom")))
    out = bench.cpu_baseline(5, "stub", 0.1)
    assert out["value"] is None and "IndexError" in out["error"]


def test_reference_arm_prints_one_line_and_other_ranks_stay_silent(monkeypatch, capsys):
    bench = _bench()
    monkeypatch.setattr(bench, "cpu_run", lambda *a, **k: {"value": 10.0, "moves": 4, "sims": 20, "seconds_max": 2.0,
                                                             "seconds_mean": 2.0, "episodes_finished": 0, "moves_per_worker": 4.0})
    a = types.SimpleNamespace(gpus=2, steps=4, warmup=1, impl="reference", games=6, sims=5, net="stub")
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(a)
    assert capsys.readouterr().out.strip() == ""
    monkeypatch.setenv("RANK", "0")
    bench.run_reference(a)
    line = json.loads(capsys.readouterr().out.strip())
    assert line["impl"] == "reference" and line["value"] == 10.0 and line["e2e"]["value"] == 10.0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["ms_per_step"] == 500.0


def test_reference_copy_is_hash_checked(tmp_path):
    """oracle/make_ref.py: the staged reference is verified against the SHA-256 manifest written at staging time; an
    edited copy is refused (bench.py then falls back to the labelled port instead of timing a modified reference)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import make_ref
    finally:
        sys.path.pop(0)
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs /root/reference (build container)")
    dest = str(tmp_path / "_ref")
    make_ref.stage("/root/reference", dest)
    assert make_ref.verify(dest)
    with open(os.path.join(dest, "MCTS.py"), "a") as f:
        f.write("\n# edited\n")
    assert not make_ref.verify(dest)
    assert not make_ref.verify(str(tmp_path / "missing"))

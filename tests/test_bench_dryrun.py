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
                                 states=0, nodes=None, cpu_baseline_seconds=0.0, no_cpu_baseline=True, no_e2e=False)
    bench.run_ours(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["metric"] == "mcts_simulations_per_sec" and line["unit"] == "sims/s" and line["n_gpus"] == 1
    assert line["steps"] == 2 and line["warmup"] == 3 and line["scaling"] == "weak" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] > 0
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

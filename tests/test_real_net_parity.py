"""north_star: "with the real net, priors and values must agree within 1e-3 (bf16 vs fp32) and root visit
distributions must agree statistically".  The priors/values half is tests/test_nnet.py; this is the visit half:
the CUDA self-play loop (fused binary16 head -> tree kernel) against the oracle MCTS driven by the SAME weights
evaluated in float32 on the CPU (NNetWrapper.predict, the reference's call), same Philox draws.

A 16-bit forward cannot reproduce a float32 search bit for bit (near-ties in PUCT flip), so the comparison is
statistical: per-game total-variation distance between the root visit distributions, agreement of the most
visited action, and the visit distribution aggregated over all games."""
import numpy as np
import pytest
import torch

from inflexion_b200.inflexion.pytorch.NNet import NNetWrapper
from inflexion_b200.selfplay import SelfPlayEngine
from oracle.inflexion_mcts import PhiloxDraws, self_play_moves


class Shape:
    board_shape, max_actions = (7, 7), 343


def sharpened_wrapper(device, seed=0, gain=40.0):
    """Random-init reference CNN whose policy head is scaled so that the priors are far from uniform (a fresh net is
    almost uniform over the valid moves, which makes every PUCT decision a near-tie and the comparison meaningless)."""
    torch.manual_seed(seed)
    w = NNetWrapper(Shape(), device=device)
    with torch.no_grad():
        w.nnet.fc3.weight.mul_(gain)
        w.nnet.fc4.weight.mul_(8.0)
    w._folded = None
    return w


@pytest.mark.gpu
def test_root_visit_distributions_agree_with_fp32_reference_search():
    G, sims, moves, seed = 32, 50, 4, 11
    gpu = sharpened_wrapper(torch.device("cuda", 0))
    cpu = sharpened_wrapper("cpu")
    cpu.nnet.load_state_dict({k: v.cpu() for k, v in gpu.nnet.state_dict().items()})
    eng = SelfPlayEngine(G, sims, gpu.predict_head, fused_head=True, temp_threshold=1000, seed=seed, record=True)
    for _ in range(moves):
        eng.play_move()
    eng.mcts.raise_on_faults()
    counts_gpu = eng.ex_counts[:moves].cpu().numpy()              # [moves][G][343]
    acts_gpu = eng.ex_actions[:moves].cpu().numpy()
    tv, same_best, agg_gpu, agg_cpu, compared = [], [], np.zeros(343), np.zeros(343), 0
    for g in range(G):
        it = self_play_moves(cpu.predict, sims, 1.0, PhiloxDraws(seed, game=g), temp_threshold=1000, max_moves=moves)
        for mv, rec in enumerate(it):
            c_ref, c_gpu = rec["counts"].astype(np.float64), counts_gpu[mv, g].astype(np.float64)
            tv.append(0.5 * np.abs(c_ref / c_ref.sum() - c_gpu / c_gpu.sum()).sum())
            same_best.append(int(np.argmax(c_ref) == np.argmax(c_gpu)))
            agg_gpu += c_gpu; agg_cpu += c_ref
            compared += 1
            if rec["action"] != int(acts_gpu[mv, g]):            # the two games diverged: later moves differ by design
                break
    tv, same_best = np.asarray(tv), np.asarray(same_best)
    agg_tv = 0.5 * np.abs(agg_gpu / agg_gpu.sum() - agg_cpu / agg_cpu.sum()).sum()
    print(f"compared {compared} searches: mean TV {tv.mean():.4f}, p90 {np.percentile(tv, 90):.4f}, identical "
          f"{np.mean(tv == 0):.2f}, same best action {same_best.mean():.2f}, aggregate TV {agg_tv:.4f}")
    assert compared >= G * 2
    assert tv.mean() < 0.05 and same_best.mean() > 0.9 and agg_tv < 0.03

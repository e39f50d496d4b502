"""BASELINE configs[2] at full size on the GPU (4 096 concurrent games x 100 sims/move), deterministic stub evaluator:
a sample of the games is compared bit for bit with the oracle, and every game is checked through size-independent
properties of the search -- conservation of visits, visits only on valid actions, a valid chosen action, determinism
in the seed, and independence of a game from the batch it is played in (the sharding property)."""
import numpy as np
import pytest
import torch

from inflexion_b200.mcts_batched import BatchedMCTS
from inflexion_b200.state import STATE_DTYPE, pack_states, unpack_masks
from inflexion_b200 import rules
from inflexion_b200.stubnet import stub_kernel_evaluator, stub_policy_value
from oracle.inflexion_mcts import PhiloxDraws, self_play_moves

G, SIMS, MOVES, SEED = 4096, 100, 4, 77


def play(be, games, tags, moves=MOVES):
    m = BatchedMCTS(len(tags), SIMS, 1.0, seed=SEED, backend=be)
    m.reset(pack_states(np.zeros((len(tags), 49), dtype=np.int64), 1, 0, 0, 343, tags=tags))
    ev = stub_kernel_evaluator(m)
    out = []
    for _ in range(moves):
        roots = m.roots().clone()
        m.search(ev)
        m.raise_on_faults()
        st = m.root_stats()
        acts = m.sample_actions(30)
        out.append((roots, st["counts"].clone(), st["ns"].clone(), acts.clone()))
        m.advance(acts)
    return out


@pytest.mark.gpu
def test_4096_games_x_100_sims_properties_and_oracle_sample(cuda_backend):
    be = cuda_backend
    full = play(be, G, np.arange(G))
    again = play(be, G, np.arange(G), moves=2)
    sub_ids = np.array([0, 1, 777, 2048, 4095, 3000, 17, 1234])
    sub = play(be, len(sub_ids), sub_ids)
    for mv, (roots, counts, ns, acts) in enumerate(full):
        masks, _, _ = rules.mask_step(roots, None, backend=be)
        valid = torch.from_numpy(unpack_masks(masks.cpu().numpy())).to(be.device).bool()
        assert bool((counts[~valid] == 0).all())                              # visits only on valid actions
        # conservation: every simulation that passes the root increments exactly one root edge and Ns (MCTS.py:136-144)
        assert torch.equal(counts.sum(dim=1), ns)
        live = roots[:, 45] == 0                      # a game can already be over at ply 3 (its only stack captured)
        assert bool(live.all()) if mv < 3 else bool(live.any())
        # the first search of an episode expands the root with its first simulation; later roots are reused subtrees
        assert bool((ns[live] == SIMS - 1).all()) if mv == 0 else bool((ns[live] >= SIMS).all())
        assert bool((ns[~live] == 0).all()) and bool((acts[~live] == -1).all())     # finished games are never searched
        a = acts.long().clamp_min(0)
        idx = torch.arange(G, device=be.device)
        assert bool((acts[live] >= 0).all()) and bool(valid[idx, a][live].all())
        assert bool((counts[idx, a][live] > 0).all())                         # sampled from the visit distribution
        if mv < 2:                                                            # determinism in the seed
            assert torch.equal(counts, again[mv][1]) and torch.equal(acts, again[mv][3])
        # a game does not depend on the batch it is played in (what sharding over ranks relies on)
        assert torch.equal(counts[torch.from_numpy(sub_ids).to(be.device)], sub[mv][1])
        assert torch.equal(acts[torch.from_numpy(sub_ids).to(be.device)], sub[mv][3])
    # bit-exact against the oracle for a sample of the 4 096 games

    def predict(planes):
        pi, v = stub_policy_value(torch.from_numpy(np.ascontiguousarray(planes))[None])
        return pi.numpy()[0], v.numpy()[0]
    for g in (0, 777, 4095):
        it = self_play_moves(predict, SIMS, 1.0, PhiloxDraws(SEED, game=g), temp_threshold=30, max_moves=MOVES)
        for mv, rec in enumerate(it):
            assert np.array_equal(full[mv][1][g].cpu().numpy(), rec["counts"]), (g, mv)
            assert int(full[mv][3][g]) == rec["action"]

"""Rules kernels (through the C ABI) against the golden transitions and the numpy oracle -- bit exact."""
import os

import numpy as np
import pytest
import torch

from inflexion_b200 import rules
from inflexion_b200.state import (STATE_DTYPE, pack_states, players_of, unpack_boards, unpack_masks)
from oracle import inflexion_rules as R


def to_np_states(t):
    return t.cpu().numpy().view(STATE_DTYPE).reshape(-1)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "rules_transitions.npz"))


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(0)
    b = rng.integers(-6, 7, size=(500, 49))
    st = pack_states(b, rng.choice([1, -1], 500), rng.integers(0, 344, 500))
    assert np.array_equal(unpack_boards(st), b)
    assert np.array_equal(st["total_power"], np.abs(b).sum(axis=1).astype(np.uint8))


def test_golden_transitions(backend, gold):
    for mt in np.unique(gold["max_turns"]):
        sel = gold["max_turns"] == mt
        st = pack_states(gold["board"][sel], gold["player"][sel], gold["turn"][sel], 0, int(mt))
        t = backend.states_tensor(st)
        acts = torch.from_numpy(gold["action"][sel].astype(np.int32)).to(backend.device)
        masks, nxt, status = rules.mask_step(t, acts, backend=backend)
        assert int(status.abs().sum()) == 0
        assert np.array_equal(np.packbits(unpack_masks(masks.cpu().numpy()), axis=1), gold["mask"][sel])
        n = to_np_states(nxt)
        assert np.array_equal(unpack_boards(n), gold["next_board"][sel])
        assert np.array_equal(n["outcome"], gold["next_outcome"][sel])
        assert np.array_equal(players_of(n), gold["next_player"][sel])
        assert np.array_equal(n["turn"], gold["next_turn"][sel])
        assert np.array_equal(n["total_power"], np.abs(gold["next_board"][sel].astype(int)).sum(axis=1))
        pl = rules.planes(t, dtype=torch.int64, backend=backend).cpu().numpy()
        assert np.array_equal(pl.reshape(-1, 4, 49), gold["planes"][sel])
        plf = rules.planes(t, dtype=torch.float32, backend=backend).cpu().numpy()
        assert np.array_equal(plf.reshape(-1, 4, 49), gold["planes"][sel].astype(np.float32))
        inf = rules.info(t, backend=backend).cpu().numpy()
        assert np.array_equal(inf[:, 0], gold["score"][sel])
        assert np.array_equal(inf[:, 2], np.abs(gold["board"][sel].astype(int)).sum(axis=1))
        assert np.array_equal(inf[:, 3], unpack_masks(masks.cpu().numpy()).sum(axis=1))


def _random_states_check(backend, n, seed):
    b, p, t, a = R.random_reachable_states(n, seed=seed)
    st = backend.states_tensor(pack_states(b, p, t))
    masks, nxt, status = rules.mask_step(st, torch.from_numpy(a.astype(np.int32)).to(backend.device), backend=backend)
    assert int(status.abs().sum()) == 0
    assert np.array_equal(unpack_masks(masks.cpu().numpy()), R.batch_valid_mask(b, p))
    nb, npl, nt, out = R.batch_step(b, p, t, a)
    n_ = to_np_states(nxt)
    assert np.array_equal(unpack_boards(n_), nb)
    assert np.array_equal(n_["outcome"], out) and np.array_equal(n_["turn"], nt) and np.array_equal(players_of(n_), npl)
    pl = rules.planes(nxt, dtype=torch.int64, backend=backend).cpu().numpy()
    assert np.array_equal(pl, R.batch_planes(nb, npl, nt))


def test_random_reachable_states_small(backend):
    _random_states_check(backend, 3000, seed=1)


def _oracle_check_transitions(states_np, acts_np, masks_np, nxt_np):
    b, p, t = unpack_boards(states_np), players_of(states_np), states_np["turn"].astype(np.int64)
    assert np.array_equal(unpack_masks(masks_np), R.batch_valid_mask(b, p))
    nb, npl, nt, out = R.batch_step(b, p, t, acts_np)
    assert np.array_equal(unpack_boards(nxt_np), nb)
    assert np.array_equal(nxt_np["outcome"], out) and np.array_equal(nxt_np["turn"], nt)
    assert np.array_equal(players_of(nxt_np), npl)
    assert np.array_equal(nxt_np["total_power"], np.abs(nb).sum(axis=1))


def _device_playouts(backend, n, seed, check_every, check_count):
    """Uniform-random valid play on the device from the start position for a uniform number of plies in
    [0, 343] (BASELINE config 2); a strided sample of the transitions is checked against the oracle on the way."""
    from oracle.inflexion_mcts import philox4x32
    rng = np.random.default_rng(seed)
    plies = torch.from_numpy(rng.integers(0, 344, size=n)).to(backend.device)
    states = backend.states_tensor(pack_states(np.zeros((n, 49), dtype=np.int64), 1, 0, 0, 343, tags=np.arange(n)))
    for t in range(343):
        acts = rules.random_actions(states, seed, t, backend=backend)
        masks, nxt, status = rules.mask_step(states, acts, backend=backend)
        if t % check_every == 0:
            idx = torch.from_numpy(rng.choice(n, size=min(check_count, n), replace=False)).to(backend.device)
            s_np, a_np = to_np_states(states[idx]), acts[idx].cpu().numpy()
            ok = a_np >= 0
            _oracle_check_transitions(s_np[ok], a_np[ok], masks[idx].cpu().numpy()[ok], to_np_states(nxt[idx])[ok])
            # the draw itself: the (w0 % count)-th valid action in ascending order
            valid = unpack_masks(masks[idx].cpu().numpy())
            for j in np.flatnonzero(ok)[:64]:
                w0 = philox4x32((int(s_np["tag"][j]), t, 0, 3), (seed & 0xFFFFFFFF, seed >> 32))[0]
                assert a_np[j] == np.flatnonzero(valid[j])[w0 % int(valid[j].sum())]
        live = (plies > t) & (status == 0) & (nxt[:, 45] == 0)      # a game that just ended stays on its last position
        states = torch.where(live[:, None], nxt, states)
    return states


def test_device_playouts_small(backend):
    states = _device_playouts(backend, 2000, seed=11, check_every=7, check_count=500)
    acts = rules.random_actions(states, 99, 0, backend=backend)
    masks, nxt, status = rules.mask_step(states, acts, backend=backend)
    assert int(status.abs().sum()) == 0
    _oracle_check_transitions(to_np_states(states), acts.cpu().numpy(), masks.cpu().numpy(), to_np_states(nxt))


@pytest.mark.gpu
def test_random_reachable_states_1m(cuda_backend):
    """BASELINE config 2: 1M random reachable states, move mask + step bit-exact against the oracle."""
    be = cuda_backend
    states = _device_playouts(be, 1_000_000, seed=0, check_every=49, check_count=20000)
    acts = rules.random_actions(states, 7, 0, backend=be)
    masks, nxt, status = rules.mask_step(states, acts, backend=be)
    assert int(status.abs().sum()) == 0
    s_np = to_np_states(states)
    assert s_np["turn"].max() >= 340 and (s_np["total_power"] > 48).any()
    _oracle_check_transitions(s_np, acts.cpu().numpy(), masks.cpu().numpy(), to_np_states(nxt))
    pl = rules.planes(nxt[:100000].contiguous(), dtype=torch.int64, backend=be).cpu().numpy()
    n_np = to_np_states(nxt[:100000])
    assert np.array_equal(pl, R.batch_planes(unpack_boards(n_np), players_of(n_np), n_np["turn"].astype(np.int64)))


def test_ragged_sizes_and_empty(backend):
    b, p, t, a = R.random_reachable_states(300, seed=5)
    for n in (0, 1, 2, 127, 128, 129, 255, 257):
        st = backend.states_tensor(pack_states(b[:n], p[:n], t[:n])) if n else backend.empty((0, 64), torch.uint8)
        acts = torch.from_numpy(a[:n].astype(np.int32)).to(backend.device)
        masks, nxt, status = rules.mask_step(st, acts, backend=backend)
        assert masks.shape == (n, 7) and nxt.shape == (n, 64)
        if n:
            nb, _, _, out = R.batch_step(b[:n], p[:n], t[:n], a[:n])
            assert np.array_equal(unpack_boards(to_np_states(nxt)), nb)
            assert np.array_equal(unpack_masks(masks.cpu().numpy()), R.batch_valid_mask(b[:n], p[:n]))
    # mask only / step only
    st = backend.states_tensor(pack_states(b, p, t))
    masks, nxt, status = rules.mask_step(st, None, backend=backend)
    assert nxt is None and np.array_equal(unpack_masks(masks.cpu().numpy()), R.batch_valid_mask(b, p))
    acts = torch.from_numpy(a.astype(np.int32)).to(backend.device)
    m2, nxt, status = rules.mask_step(st, acts, want_masks=False, backend=backend)
    assert m2 is None and np.array_equal(unpack_boards(to_np_states(nxt)), R.batch_step(b, p, t, a)[0])


def test_invalid_actions_are_flagged_not_applied(backend):
    b, p, t, a = R.random_reachable_states(200, seed=9)
    mask = R.batch_valid_mask(b, p)
    bad = np.array([int(np.nonzero(m == 0)[0][i % (m == 0).sum()]) for i, m in enumerate(mask)], dtype=np.int32)
    bad[::7] = 343
    bad[1::7] = -1
    st = backend.states_tensor(pack_states(b, p, t))
    _, nxt, status = rules.mask_step(st, torch.from_numpy(bad).to(backend.device), backend=backend)
    assert (status.cpu().numpy() == 16).all()
    assert np.array_equal(nxt.cpu().numpy(), st.cpu().numpy())
    # a finished game accepts no move
    over = pack_states(b[:4], p[:4], t[:4], outcomes=[1, 2, 3, 2])
    _, nxt, status = rules.mask_step(backend.states_tensor(over), torch.from_numpy(a[:4].astype(np.int32)).to(backend.device),
                                     backend=backend)
    assert (status.cpu().numpy() == 16).all()


def test_symmetries_and_random_symmetry(backend, golden_dir):
    g = np.load(os.path.join(golden_dir, "symmetries.npz"))
    planes = np.arange(4 * 49).reshape(1, 4, 7, 7)
    policy = np.arange(7 * 49).reshape(1, 7, 7, 7)
    for arr, key in ((planes, "planes36"), (policy, "policy36")):
        for dt in (np.int64, np.float32, np.int32, np.float64):
            out = rules.symmetries(torch.from_numpy(arr.astype(dt)).to(backend.device), backend=backend).cpu().numpy()
            assert np.array_equal(out[0], g[key].astype(dt))
    # planes under every (k, shift, axis) of random_symmetry
    b, p, t, _ = R.random_reachable_states(126, seed=2)
    st = backend.states_tensor(pack_states(b, p, t))
    sym = np.arange(126, dtype=np.int32)
    out = rules.planes(st, torch.from_numpy(sym).to(backend.device), dtype=torch.int64, backend=backend).cpu().numpy()
    ref = R.batch_planes(b, p, t)
    for i in range(126):
        k, sh, ax = i % 6, (i // 6) % 7, i // 42
        assert np.array_equal(out[i], R.random_symmetry(ref[i], k, sh, ax)), (k, sh, ax)


def test_host_buffer_entry_points(backend):
    b, p, t, a = R.random_reachable_states(500, seed=4)
    st = pack_states(b, p, t)
    masks, nxt, status = rules.mask_step_host(st, a, backend=backend)
    assert np.array_equal(unpack_masks(masks), R.batch_valid_mask(b, p)) and not status.any()
    nb, npl, nt, out = R.batch_step(b, p, t, a)
    assert np.array_equal(unpack_boards(nxt), nb) and np.array_equal(nxt["outcome"], out)
    assert np.array_equal(rules.planes_host(nxt, backend=backend), R.batch_planes(nb, npl, nt))
    x = np.arange(2 * 7 * 49, dtype=np.float32).reshape(2, 7, 7, 7)
    sy = rules.symmetries_host(x, backend=backend)
    assert np.array_equal(sy[1], np.array(R.symmetries(x[1])))
    assert np.array_equal(rules.info_host(st, backend=backend)[:, 3], R.batch_valid_mask(b, p).sum(axis=1))

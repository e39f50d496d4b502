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


@pytest.mark.gpu
def test_random_reachable_states_1m(cuda_backend):
    """BASELINE config 2: 1M random reachable states, move mask + step bit-exact."""
    _random_states_check(cuda_backend, 1_000_000, seed=0)


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

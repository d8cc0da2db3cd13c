"""GPU parity of Philox lockstep steps at awkward batch sizes and with every kind of no-op mixed in -- written for the
pooled common-case step kernel (pool_step_kernel, a measured and rejected variant that is compiled only with
-DECG_POOL=1: ECG_LIB=<that build> runs these tests through it, ECG_POOL_MIN_BOARDS=0 sends every batch its way), and
kept because the default two-kernel step has to pass the same cases.
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.oracle import Oracle, ST_CASCADE_CAP  # noqa: E402
from test_gpu_parity import KEY, cfg_of, np_  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ecg_b200
    return ecg_b200


@pytest.fixture()
def pool_always():
    old = os.environ.get("ECG_POOL_MIN_BOARDS")
    os.environ["ECG_POOL_MIN_BOARDS"] = "0"
    yield
    if old is None:
        del os.environ["ECG_POOL_MIN_BOARDS"]
    else:
        os.environ["ECG_POOL_MIN_BOARDS"] = old


@pytest.mark.parametrize("n", [1, 31, 257, 5000, 40001])
@pytest.mark.parametrize("types", [6, 8])
def test_pool_philox_episodes_vs_oracle(E, pool_always, n, types):
    shape, moves, board0 = (9, 9, types), 12, 999
    o = Oracle(*shape)
    bb = E.BatchedBoards(cfg_of(E, shape), n, moves, key=KEY, board0=board0)
    init = np_(bb.array)
    final, total, steps = o.philox_episode_batch(init, KEY, board0, moves)
    for t in range(moves):
        bb.apply_action(None)
    assert np.array_equal(np_(bb.array), final)
    assert np.array_equal(np_(bb.reward), total)
    assert np.array_equal(np_(bb.legal_mask()), o.legal_mask_batch(final))


@pytest.mark.parametrize("shape", [(9, 9, 6), (9, 9, 3), (9, 9, 8), (9, 9, 2)])
def test_pool_fuzz_dense_boards(E, pool_always, shape):
    """given actions (legal and illegal), planted specials and empty cells: every output of the step"""
    rng = np.random.default_rng(shape[2] + 41)
    o = Oracle(*shape)
    R, Cc, T = shape
    n = 6000 if T > 2 else 400  # two types cascade to the cap (1024 iterations per board)
    sp = [o.cfg.h_line, o.cfg.v_line, o.cfg.bomb, o.cfg.mega_token]
    tl = rng.integers(2, T + 1, size=n)
    b = np.stack([rng.integers(1, t + 1, size=(R, Cc)) for t in tl]).astype(np.int64)
    for i in range(0, n, 3):
        for _ in range(int(rng.integers(0, 4))):
            b[i, rng.integers(R), rng.integers(Cc)] = sp[rng.integers(4)]
        if i % 2 == 0:
            b[i, rng.integers(R), rng.integers(Cc)] = 0
    lo = o.legal_mask_batch(b)
    acts = rng.integers(-1, o.A + 1, size=n)
    for i in range(0, n, 2):
        la = np.flatnonzero(lo[i])
        if len(la):
            acts[i] = la[rng.integers(len(la))]
    moves = rng.integers(0, 3, size=n).astype(np.int32)
    import torch
    bb = E.BatchedBoards(cfg_of(E, shape), n, arrays=b, key=KEY, board0=77)
    bb.moves_left.copy_(torch.from_numpy(moves))
    bb.step_ctr = 5
    bb.apply_action(acts.astype(np.int32))
    live = (moves >= 1) & (acts >= 0) & (acts < o.A)
    # the oracle steps board i of a batch with Philox id board0 + i: feed it the whole batch and compare the live rows
    ro = o.step_batch(b, np.where(live, acts, 0), mode="philox", key=KEY, board0=77, step_ctr=5)
    got = np_(bb.array)
    assert np.array_equal(got[live], ro["boards"][live])
    assert np.array_equal(got[~live], b[~live])
    assert np.array_equal(np_(bb.step_reward)[live], ro["reward"][live]) and not np_(bb.step_reward)[~live].any()
    assert np.array_equal(np_(bb.cascades)[live], ro["cascades"][live])
    assert np.array_equal(np_(bb.status)[live], ro["status"][live])
    st = np_(bb.status)
    assert (st[moves < 1] == E.ST_TERMINAL).all()
    assert (st[(moves >= 1) & ~live] == E.ST_BAD_ACTION).all()
    assert np.array_equal(np_(bb.moves_left), moves - live)
    assert np.array_equal(np_(bb.legal_mask()), o.legal_mask_batch(got))
    if T == 2:
        assert (st[live] & ST_CASCADE_CAP).any()


def test_pool_matches_lane_kernel_at_size(E):
    """a batch large enough to take the pooled kernel (in a build that has it) against the same batch forced through the
    lane kernel: boards, rewards, cascades, status, masks, scores of every step identical"""
    n, steps = 1 << 21, 4
    a = E.BatchedBoards(cfg_of(E, (9, 9, 6)), n, 30, key=KEY, board0=5)
    b = a.clone()
    import torch
    for t in range(steps):
        a.apply_action(None)
        os.environ["ECG_POOL_MIN_BOARDS"] = str(1 << 40)
        try:
            b.apply_action(None)
        finally:
            del os.environ["ECG_POOL_MIN_BOARDS"]
        for name in ("boards", "mask", "score", "step_reward", "cascades", "status", "flags", "moves_left", "last_actions"):
            assert torch.equal(getattr(a, name), getattr(b, name)), (name, t)

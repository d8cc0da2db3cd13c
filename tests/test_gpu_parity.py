"""GPU parity: the CUDA path (through the Python host layer -> ctypes -> C-ABI of libecg.so) against the CPU
oracle and the reference-generated golden vectors.  Everything here is integer work: the bar is bit-exact.
Run on the B200 box:  python -m pytest tests -m gpu -x -q"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ALL_SHAPES, GOLDEN, SHAPES  # noqa: E402
from oracle.oracle import Oracle, ST_CASCADE_CAP, ST_SHUFFLE_CAP, ST_STREAM_OVERFLOW  # noqa: E402

pytestmark = pytest.mark.gpu
KEY = 0x1234567890ABCDEF


@pytest.fixture(scope="module")
def E():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ecg_b200
    return ecg_b200


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def cfg_of(E, shape, seed=1):
    return E.BoardConfig(seed=seed, rows=shape[0], columns=shape[1], types=shape[2])


def np_(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_pack_unpack_roundtrip_and_bad_cells(E, shape):
    import torch
    o = Oracle(*shape)
    rng = np.random.default_rng(1)
    n = 1000 + 7  # not a multiple of the 32-board tile
    vals = np.array(list(range(0, min(o.cfg.type_mask, 11) + 1)) + [o.cfg.h_line, o.cfg.v_line, o.cfg.bomb, o.cfg.mega_token])
    b = vals[rng.integers(len(vals), size=(n, shape[0], shape[1]))].astype(np.int64)
    bb = E.BatchedBoards(cfg_of(E, shape), n, arrays=b, key=KEY)
    assert np.array_equal(np_(bb.array), b)
    assert np.array_equal(np_(bb.observe(torch.uint8)), b.astype(np.uint8))
    bb2 = E.BatchedBoards(cfg_of(E, shape), n, arrays=torch.as_tensor(b.astype(np.uint8)), key=KEY)
    assert np.array_equal(np_(bb2.array), b)
    bad = b.copy()
    bad[5, 0, 0] = o.cfg.h_line + 1  # typed special: outside the engine's closed code set
    with pytest.raises(ValueError):
        E.BatchedBoards(cfg_of(E, shape), n, arrays=bad, key=KEY)


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_legal_mask_golden(E, shape):
    d = load("funcs_%dx%dx%d.npz" % shape)
    o = Oracle(*shape)
    b = d["boards"].astype(np.int64)
    legal = np.unpackbits(d["legal"], axis=-1)[..., :o.A].astype(bool)
    bb = E.BatchedBoards(cfg_of(E, shape), len(b), arrays=b, key=KEY)
    assert np.array_equal(np_(bb.legal_mask()), legal)
    assert bb.legal_actions[3] == np.flatnonzero(legal[3]).tolist()


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_golden_single_steps_replay(E, shape):
    d = load("steps_%dx%dx%d.npz" % shape)
    n = len(d["actions"])
    bb = E.BatchedBoards(cfg_of(E, shape), n, arrays=d["before"].astype(np.int64), refill="replay",
                         seeds=d["seeds"].astype(np.int64), stream_len=4096)
    bb.apply_action(d["actions"].astype(np.int32))
    assert not np_(bb.status).any()
    assert np.array_equal(np_(bb.array), d["after"])
    assert np.array_equal(np_(bb.step_reward), d["rewards"])
    assert np.array_equal(np_(bb.cascades), d["cascades"])
    o = Oracle(*shape)
    assert np.array_equal(np_(bb.legal_mask()), o.legal_mask_batch(d["after"].astype(np.int64)))


def test_golden_shuffle_cases(E):
    d = load("shuffle.npz")
    for i in range(len(d["actions"])):
        R, Cc, T = (int(x) for x in d["shape"][i])
        bb = E.BatchedBoards(cfg_of(E, (R, Cc, T)), 1, arrays=d["before"][i:i + 1, :R, :Cc].astype(np.int64),
                             refill="replay", seeds=[int(d["seeds"][i])])
        bb.apply_action([int(d["actions"][i])])
        assert int(bb.status[0]) == 0
        assert np.array_equal(np_(bb.array)[0], d["after"][i, :R, :Cc])
        assert int(bb.step_reward[0]) == d["rewards"][i] and int(bb.cascades[0]) == d["cascades"][i]


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_golden_episodes_random_task(E, shape):
    """samplerTasks.random_task (:9-14) replayed on the GPU: init boards from the MT19937 stream, picks by
    numpy's masked rejection on the same stream, 20 steps; then the same episodes in ONE rollout kernel."""
    d = load("episodes_%dx%dx%d.npz" % shape)
    seeds = d["seeds"].astype(np.int64)
    moves = int(d["moves"])
    mk = lambda: E.BatchedBoards(cfg_of(E, shape), len(seeds), moves, refill="replay", seeds=seeds, stream_len=8192)  # noqa: E731
    bb = mk()
    assert np.array_equal(np_(bb.array), d["init"])
    for t in range(moves):
        legal = np.unpackbits(d["legal"][:, t], axis=-1)[..., :bb.cfg.action_space].astype(bool)
        assert np.array_equal(np_(bb.legal_mask()), legal)
        bb.apply_action(None)
        assert np.array_equal(np_(bb.last_actions), d["actions"][:, t])
        assert np.array_equal(np_(bb.array), d["boards"][:, t])
        assert np.array_equal(np_(bb.step_reward), d["rewards"][:, t])
        assert np.array_equal(np_(bb.cascades), d["cascades"][:, t])
    assert np.array_equal(np_(bb.reward), d["rewards"].sum(axis=1))
    assert bool(bb.is_terminal.all())
    rb = mk()
    total = rb.rollout()
    assert np.array_equal(np_(total), d["rewards"].sum(axis=1))
    assert np.array_equal(np_(rb.array), d["boards"][:, -1])


def test_config2_4096_lockstep_boards_vs_oracle(E):
    """BASELINE.json configs[1]: 4096 lockstep 9x9x6 boards, seeds 1..4096, random legal actions,
    step + cascade + reward, bit-exact against the reference semantics (oracle, itself pinned to the reference)."""
    import torch
    n, moves = 4096, 20
    o = Oracle(9, 9, 6)
    seeds = np.arange(1, n + 1, dtype=np.int64)
    env = E.BatchedMatch3Env(n, seed=1, refill="replay", stream_len=4096, obs_dtype=torch.int64)
    assert np.array_equal(np.asarray(env.board.seeds), seeds)
    boards = np_(env.init())
    raw = np.stack([Oracle.mt_raw(int(s), 4096) for s in seeds])
    score = np.zeros(n, dtype=np.int64)
    for t in range(moves):
        acts = np_(env.board.random_action()) if t % 2 else None  # both pick paths: separate kernel / fused
        pos_before = np_(env.board.stream_pos).copy()
        obs, rew, done, won, _ = env.step(acts)
        a = np_(env.board.last_actions)
        legal = o.legal_mask_batch(boards)
        assert legal[np.arange(n), a].all()
        res = o.step_batch(boards, a, mode="replay", raw=raw)
        assert not res["status"].any() and not np_(env.board.status).any()
        assert np.array_equal(np_(obs), res["boards"])
        assert np.array_equal(np_(rew), res["reward"])
        assert np.array_equal(np_(env.board.cascades), res["cascades"])
        assert np.array_equal(np_(env.board.legal_mask()), res["legal"])
        score += res["reward"]
        assert np.array_equal(np_(won), score >= 500)
        assert np.array_equal(np_(done), (score >= 500) | (t == moves - 1))
        boards = res["boards"]
    total, steps = o.random_episode_batch(seeds.astype(np.uint32), moves)
    assert np.array_equal(score, total)
    st = E.dist.stats_dict(env.board.episode_stats())
    assert st["episodes"] == n and st["min"] == score.min() and st["max"] == score.max()
    assert abs(st["mean"] - score.mean()) < 1e-9 and st["wins"] == int((score >= 500).sum())


@pytest.mark.parametrize("shape", [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8)])
def test_philox_episodes_vs_oracle(E, shape):
    """config 3/4 semantics at a size the oracle finishes in seconds: Philox-keyed init, picks and refills."""
    n, moves, board0 = 8192 if shape[0] <= 9 else 2048, 20, 12345
    o = Oracle(*shape)
    bb = E.BatchedBoards(cfg_of(E, shape), n, moves, key=KEY, board0=board0)
    init = np_(bb.array)
    for i in range(8):
        assert np.array_equal(o.init_board(o.rng_philox(KEY, board0 + i, 0xFFFFFFFF)), init[i])
    final, total, steps = o.philox_episode_batch(init, KEY, board0, moves)
    lock = bb.clone()
    for t in range(moves):
        lock.apply_action(None)
    assert np.array_equal(np_(lock.array), final)
    assert np.array_equal(np_(lock.reward), total)
    roll = bb.clone()
    tot = roll.rollout()
    assert np.array_equal(np_(tot), total) and np.array_equal(np_(roll.array), final)
    # sharding: the same boards stepped as two shards with their global offsets give the same result
    half = n // 2
    a = E.BatchedBoards(cfg_of(E, shape), half, moves, key=KEY, board0=board0)
    b = E.BatchedBoards(cfg_of(E, shape), n - half, moves, key=KEY, board0=board0 + half)
    ta, tb = a.rollout(), b.rollout()
    assert np.array_equal(np.concatenate([np_(ta), np_(tb)]), total)
    assert np.array_equal(np.concatenate([np_(a.array), np_(b.array)]), final)


@pytest.mark.parametrize("shape", [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8), (6, 6, 3), (5, 5, 2), (7, 7, 5),
                                   (9, 9, 3), (9, 9, 8), (6, 6, 11)])
def test_fuzz_dense_boards_philox(E, shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[2] + 1)
    o = Oracle(*shape)
    R, Cc, T = shape
    n = 3000
    sp = [o.cfg.h_line, o.cfg.v_line, o.cfg.bomb, o.cfg.mega_token]
    tl = rng.integers(2, T + 1, size=n)
    b = np.stack([rng.integers(1, t + 1, size=(R, Cc)) for t in tl]).astype(np.int64)
    for i in range(0, n, 3):
        for _ in range(int(rng.integers(0, 4))):
            b[i, rng.integers(R), rng.integers(Cc)] = sp[rng.integers(4)]
        if i % 2 == 0:
            b[i, rng.integers(R), rng.integers(Cc)] = 0
    lo = o.legal_mask_batch(b)
    acts = rng.integers(0, o.A, size=n)
    for i in range(0, n, 2):
        la = np.flatnonzero(lo[i])
        if len(la):
            acts[i] = la[rng.integers(len(la))]
    for i in range(0, n, 5):
        (r1, c1), (r2, c2) = o.decode(int(acts[i]))
        b[i, r1, c1] = sp[rng.integers(4)]
        if rng.integers(2):
            b[i, r2, c2] = sp[rng.integers(4)]
    bb = E.BatchedBoards(cfg_of(E, shape), n, arrays=b, key=KEY, board0=77)
    assert np.array_equal(np_(bb.legal_mask()), o.legal_mask_batch(b))
    bb.step_ctr = 5
    bb.apply_action(acts.astype(np.int32))
    ro = o.step_batch(b, acts, mode="philox", key=KEY, board0=77, step_ctr=5)
    assert np.array_equal(np_(bb.array), ro["boards"])
    assert np.array_equal(np_(bb.step_reward), ro["reward"])
    assert np.array_equal(np_(bb.cascades), ro["cascades"])
    assert np.array_equal(np_(bb.status), ro["status"])
    assert np.array_equal(np_(bb.legal_mask()), ro["legal"])


def test_caps_flags_and_noops(E):
    import torch
    rng = np.random.default_rng(5)
    o = Oracle(9, 9, 2)
    b = rng.integers(1, 3, size=(64, 9, 9)).astype(np.int64)
    acts = rng.integers(0, o.A, size=64)
    bb = E.BatchedBoards(cfg_of(E, (9, 9, 2)), 64, arrays=b, key=KEY)
    bb.apply_action(acts.astype(np.int32))
    ro = o.step_batch(b, acts, mode="philox", key=KEY)
    assert (np_(bb.status) & ST_CASCADE_CAP).all()
    assert np.array_equal(np_(bb.array), ro["boards"]) and np.array_equal(np_(bb.step_reward), ro["reward"])
    # shuffle cap on the board that hangs the reference
    arr = np.fromfunction(lambda r, c: ((c + 2 * (r % 3)) % 6) + 1, (9, 9), dtype=np.int64).astype(np.int64)[None]
    o = Oracle(9, 9, 6)
    bb = E.BatchedBoards(cfg_of(E, (9, 9, 6)), 1, arrays=arr, refill="replay", seeds=[7])
    bb.apply_action([7])
    ro = o.step_batch(arr, [7], mode="replay", raw=Oracle.mt_raw(7, 4096))
    assert int(bb.status[0]) & ST_SHUFFLE_CAP and ro["status"][0] & ST_SHUFFLE_CAP
    assert np.array_equal(np_(bb.array), ro["boards"]) and int(bb.step_reward[0]) == ro["reward"][0]
    # terminal boards and out-of-range actions are no-ops with a flag
    b = rng.integers(1, 7, size=(4, 9, 9)).astype(np.int64)
    bb = E.BatchedBoards(cfg_of(E, (9, 9, 6)), 4, arrays=b, key=KEY)
    bb.moves_left.copy_(torch.tensor([0, 1, 5, 5], dtype=torch.int32))
    bb.apply_action([3, 3, -1, 144])
    st = np_(bb.status)
    assert st[0] == E.ST_TERMINAL and st[2] == E.ST_BAD_ACTION and st[3] == E.ST_BAD_ACTION and st[1] == 0
    after = np_(bb.array)
    assert np.array_equal(after[[0, 2, 3]], b[[0, 2, 3]]) and not np_(bb.step_reward)[[0, 2, 3]].any()
    assert np_(bb.moves_left).tolist() == [0, 0, 5, 5]
    assert np.array_equal(np_(bb.legal_mask()), o.legal_mask_batch(after))
    # replay stream too short
    b = rng.integers(1, 3, size=(2, 9, 9)).astype(np.int64)
    bb = E.BatchedBoards(cfg_of(E, (9, 9, 6)), 2, arrays=b, refill="replay", seeds=[3, 4], stream_len=2)
    bb.apply_action([0, 1])
    assert (np_(bb.status) & ST_STREAM_OVERFLOW).all()


def test_mt19937_stream_kernel_matches_numpy(E):
    d = load("rng.npz")
    seeds = d["seeds"].astype(np.int64)
    bb = E.BatchedBoards(cfg_of(E, (9, 9, 6)), len(seeds), refill="replay", seeds=seeds, stream_len=700)
    got = np_(bb.stream).view(np.uint32).reshape(len(seeds), 700)
    assert np.array_equal(got, d["raw"])


def test_match3env_dropin_contract(E):
    """README usage: env.reset() / env.board.random_action() / env.step(a) with reference return types."""
    d = load("episodes_9x9x6.npz")
    e = 4
    seed = int(d["seeds"][e])
    env = E.Match3Env(seed=seed)
    obs, info = env.reset()
    assert obs.dtype == np.int64 and obs.shape == (9, 9) and info == {}
    assert np.array_equal(obs, d["init"][e])
    score = 0
    for t in range(20):
        assert env.board.legal_actions == np.flatnonzero(np.unpackbits(d["legal"][e, t])[:144]).tolist()
        a = int(d["actions"][e, t])
        obs, reward, done, won, info = env.step(a)
        assert isinstance(reward, int) and isinstance(done, bool) and isinstance(won, bool)
        assert np.array_equal(obs, d["boards"][e, t]) and reward == d["rewards"][e, t]
        score += reward
        assert env.score == score and env.moves_taken == t + 1
        assert won == (score >= 500) and done == (won or t == 19)
    a = env.board.random_action()
    assert a == -1 or a in env.board.legal_actions
    obs2, _ = env.reset()
    assert np.array_equal(obs2, d["init"][e])  # env.py:62: reset() without a seed keeps the seed


@pytest.mark.parametrize("obs_format", ["uint8", "nibbles", "uint8+host_expand"])
def test_host_stepper_matches_device_api(E, obs_format):
    """HostStepper (host buffers in pinned memory, chunked over CUDA streams: the call bench.py's e2e times) returns
    the same observations, rewards and done / won flags as BatchedMatch3Env.step on a twin environment, and both
    match the oracle (Philox mode, ragged chunk sizes).  obs_format="nibbles" carries the same information in half
    the PCIe bytes (4-bit cell codes, int16 reward and actions); host_expand=True returns the uint8 form, byte for
    byte, from 4-bit codes widened by the library's host thread pool."""
    import torch
    expand = obs_format.endswith("+host_expand")
    obs_format = obs_format.split("+")[0]
    n, moves = 3000 + (1 if obs_format == "nibbles" or expand else 0), 6  # not a multiple of the chunk count or of the tile
    o = Oracle(9, 9, 6)
    a_env = E.BatchedMatch3Env(n, seed=11, num_moves=moves, env_goal=60)
    b_env = E.BatchedMatch3Env(n, seed=11, num_moves=moves, env_goal=60)
    boards = np_(a_env.init()).astype(np.int64)
    assert np.array_equal(boards, np_(b_env.init()))
    hs = E.HostStepper(a_env, chunks=5, obs_format=obs_format, host_expand=expand, expand_threads=3 if expand else None)
    small = torch.int16 if obs_format == "nibbles" else torch.int32
    assert hs.d2h_bytes == n * ((41 + 2 + 2) if obs_format == "nibbles" else (41 + 4 + 2) if expand else (81 + 4 + 2))
    score = np.zeros(n, dtype=np.int64)
    for t in range(moves):
        acts = hs.random_action()
        assert acts.is_pinned() and acts.dtype == small
        a = acts.numpy().astype(np.int32)
        obs, rew, done, won, _ = hs.step(acts)
        for x in (obs, rew, done, won):
            assert x.device.type == "cpu" and x.is_pinned()
        assert done.dtype == torch.bool and won.dtype == torch.bool and rew.dtype == small
        if obs_format == "nibbles":
            assert obs.shape == (n, 41) and obs.dtype == torch.uint8
            cells = hs.decode_obs(obs)
        else:
            cells = obs.numpy()
        obs2, rew2, done2, won2, _ = b_env.step(torch.from_numpy(a).to(b_env.board.device))
        assert np.array_equal(cells, np_(obs2)) and np.array_equal(rew.numpy(), np_(rew2))
        assert np.array_equal(done.numpy(), np_(done2)) and np.array_equal(won.numpy(), np_(won2))
        res = o.step_batch(boards, a, mode="philox", key=a_env.board.key, board0=0, step_ctr=t)
        assert np.array_equal(cells, res["boards"]) and np.array_equal(rew.numpy(), res["reward"])
        score += res["reward"]
        assert np.array_equal(won.numpy(), score >= 60)
        assert np.array_equal(done.numpy(), (score >= 60) | (t == moves - 1))
        boards = res["boards"]


@pytest.mark.parametrize("shape", [(4, 4, 4), (5, 5, 4), (6, 6, 4), (11, 11, 7), (16, 16, 8)])
def test_host_expand_other_board_sizes(E, shape):
    """host_expand on boards below the 16-byte SIMD block (scalar table), with odd and even cell counts, and on the
    largest board: the widened observation equals the device's own uint8 unpack, rewards and flags equal the twin's"""
    import torch
    n, moves = 2000 + 37, 3
    a_env = E.BatchedMatch3Env(n, shape[1], shape[0], shape[2], seed=5, num_moves=moves, env_goal=40)
    b_env = E.BatchedMatch3Env(n, shape[1], shape[0], shape[2], seed=5, num_moves=moves, env_goal=40)
    a_env.init()
    b_env.init()
    hs = E.HostStepper(a_env, chunks=3, host_expand=True, expand_threads=4, expand_pieces=3)
    try:
        for t in range(moves):
            acts = hs.random_action()
            obs, rew, done, won, _ = hs.step(acts)
            obs2, rew2, done2, won2, _ = b_env.step(acts.to(b_env.board.device))
            assert obs.shape == (n, shape[0], shape[1]) and obs.dtype == torch.uint8
            assert np.array_equal(obs.numpy(), np_(obs2)) and np.array_equal(rew.numpy(), np_(rew2))
            assert np.array_equal(done.numpy(), np_(done2)) and np.array_equal(won.numpy(), np_(won2))
            assert np.array_equal(obs.numpy(), np_(a_env.board.array).astype(np.uint8))
    finally:
        hs.close()


def test_host_stepper_replay_chunks(E):
    """HostStepper over a replay-mode env: every chunk addresses its own slice of the per-board MT19937 streams, tile
    tables and stream positions (two-kernel replay step per chunk); a twin env stepped in one piece must agree."""
    import torch
    n, moves = 1000 + 9, 5
    a_env = E.BatchedMatch3Env(n, seed=21, num_moves=moves, env_goal=80, refill="replay", stream_len=1024)
    b_env = E.BatchedMatch3Env(n, seed=21, num_moves=moves, env_goal=80, refill="replay", stream_len=1024)
    assert a_env.board.stream_stride == 1024 and a_env.board.tiles is not None
    hs = E.HostStepper(a_env, chunks=7)
    for t in range(moves):
        acts = hs.random_action()
        a = acts.numpy().copy()
        assert np.array_equal(a, np_(b_env.board.random_action()))
        obs, rew, done, won, _ = hs.step(acts)
        obs2, rew2, done2, won2, _ = b_env.step(torch.from_numpy(a).to(b_env.board.device))
        assert np.array_equal(obs.numpy(), np_(obs2)) and np.array_equal(rew.numpy(), np_(rew2))
        assert np.array_equal(done.numpy(), np_(done2)) and np.array_equal(won.numpy(), np_(won2))
        assert np.array_equal(np_(a_env.board.stream_pos), np_(b_env.board.stream_pos))
        assert not np_(a_env.board.status).any()


def test_nibble_observation_all_codes_and_shapes(E):
    """ecg_unpack_nibbles: 4-bit cell codes, two per byte, ceil(R*C/2) bytes per board, ragged tiles, odd R*C"""
    import ctypes as C
    import torch
    for shape in ((9, 9, 6), (5, 5, 2), (16, 16, 8), (6, 6, 4)):
        o = Oracle(*shape)
        rng = np.random.default_rng(4)
        vals = np.array(list(range(0, shape[2] + 1)) + [o.cfg.h_line, o.cfg.v_line, o.cfg.bomb, o.cfg.mega_token])
        codes = np.array(list(range(0, shape[2] + 1)) + [12, 13, 14, 15])
        for n in (1, 31, 32, 33, 1001):
            pick = rng.integers(len(vals), size=(n, shape[0], shape[1]))
            bb = E.BatchedBoards(cfg_of(E, shape), n, arrays=vals[pick].astype(np.int64), key=KEY)
            nby = (shape[0] * shape[1] + 1) // 2
            buf = torch.full((n * nby + 64,), 0xAB, dtype=torch.uint8, device=bb.device)
            N = E._native
            N.check(N.lib().ecg_unpack_nibbles(C.byref(bb.nat), C.c_void_p(bb.boards.data_ptr()),
                                               C.c_void_p(buf.data_ptr()), n, None), "ecg_unpack_nibbles")
            torch.cuda.synchronize()
            got = np_(buf[:n * nby]).reshape(n, nby)
            flat = codes[pick].reshape(n, -1)
            if flat.shape[1] % 2:
                flat = np.concatenate([flat, np.zeros((n, 1), dtype=flat.dtype)], axis=1)
            want = (flat[:, 0::2] | (flat[:, 1::2] << 4)).astype(np.uint8)
            assert np.array_equal(got, want), (shape, n)
            assert bool((buf[n * nby:] == 0xAB).all())


def test_boardv2_view_matches_golden_and_state_abc(E):
    d = load("episodes_6x6x4.npz")
    e = 2
    cfg = E.BoardConfig(seed=int(d["seeds"][e]), rows=6, columns=6, types=4)
    st = E.BoardV2(int(d["moves"]), cfg)
    assert np.array_equal(st.array, d["init"][e])
    total = 0
    for t in range(int(d["moves"])):
        assert st.legal_actions == np.flatnonzero(np.unpackbits(d["legal"][e, t])[:cfg.action_space]).tolist()
        nxt = st.apply_action(int(d["actions"][e, t]))
        assert np.array_equal(st.array, d["init"][e] if t == 0 else d["boards"][e, t - 1])  # never mutates
        total += int(d["rewards"][e, t])
        assert nxt.reward == total and nxt.n_actions == st.n_actions - 1
        assert np.array_equal(nxt.array, d["boards"][e, t])
        st = nxt
    assert st.is_terminal and st.apply_action(0) is st
    c = st.clone()
    assert c.reward == st.reward and np.array_equal(c.array, st.array)
    with pytest.raises(KeyError):
        E.BoardV2(3, cfg).apply_action(cfg.action_space)
    # greedy_action: first maximum over legal actions of the one-step reward (boardv2.py:209-218)
    s0 = E.BoardV2(5, cfg)
    o = Oracle(6, 6, 4)
    best, hi = None, -1
    for a in s0.legal_actions:
        _, r, _, _, _ = o.apply_action(o.rng_mt(cfg.seed), s0.array, a)
        if r > hi:
            hi, best = r, a
    assert s0.greedy_action == best
    bb = E.BatchedBoards(cfg, 1, 5, refill="replay", seeds=[cfg.seed])
    assert int(bb.greedy_action()[0]) == best


def test_full_size_properties(E):
    """BASELINE.json configs[2] size (2^24 boards): properties that do not need the oracle at full size,
    plus a strided sample checked against it."""
    import torch
    n, moves = 1 << 24, 3
    shape = (9, 9, 6)
    o = Oracle(*shape)
    bb = E.BatchedBoards(cfg_of(E, shape), n, moves, key=KEY)
    idx = torch.arange(0, n, 4099, device=bb.device)
    before = np_(bb.array[idx])
    acc = torch.zeros(n, dtype=torch.int64, device=bb.device)
    sample = before
    for t in range(moves):
        bb.apply_action(None)
        acc += bb.step_reward
        obs = bb.observe(torch.uint8)
        assert int(obs.min()) >= 1                       # refilled: no empty cell survives a step
        assert int(obs.max()) <= 32                      # closed value set {1..6, 8, 16, 24, 32}
        st = bb.status
        stuck = (st & E.ST_NO_LEGAL) != 0   # a random initial board without any legal move: no-op + flag
        assert not bool((st & ~(E.ST_NO_LEGAL | E.ST_SHUFFLE_CAP)).any())
        assert int(stuck.sum()) < n // 10000
        assert int(bb.cascades[~stuck].min()) >= 1 and int(bb.step_reward.min()) >= 0
        a = np_(bb.last_actions[idx])
        res = o.step_batch(sample, a, mode="philox", key=KEY, board0=0, step_ctr=t)
        # board0 + i with i the *sample* index differs from the global index: check each sampled board on its own
        for j, g in enumerate(np_(idx)[:64]):
            r1 = o.step_batch(sample[j:j + 1], a[j:j + 1], mode="philox", key=KEY, board0=int(g), step_ctr=t)
            assert np.array_equal(r1["boards"][0], np_(bb.array[idx[j]:idx[j] + 1])[0])
            assert r1["reward"][0] == int(bb.step_reward[idx[j]])
        sample = np_(bb.array[idx])
        del res
    assert torch.equal(acc, bb.reward)                   # score is the sum of the step rewards
    assert bool((bb.is_terminal | ((bb.status & E.ST_NO_LEGAL) != 0)).all())
    # a stable board has no run of three: stepping the legal mask kernel twice is idempotent
    m1 = bb.packed_mask().clone()
    bb._mask_valid = False
    assert torch.equal(m1, bb.packed_mask())


def test_full_size_replay_properties(E):
    """The reference's dynamics at BASELINE configs[2] size: 2^24 boards, two-kernel replay step, 4096 distinct MT19937
    streams; a strided sample is checked against the oracle (boards, rewards, np.random's position), the rest through
    size-independent properties."""
    import torch
    n, streams, moves = 1 << 24, 4096, 2
    shape = (9, 9, 6)
    o = Oracle(*shape)
    src = E.BatchedBoards(cfg_of(E, shape), n, 9, key=KEY)
    dev = src.device
    bb = E.BatchedBoards(cfg_of(E, shape), n, 9, refill="replay", stream_len=1024,
                         seeds=[1000 + i for i in range(streams)],
                         stream_index=torch.arange(n, device=dev).remainder(streams))
    assert bb.two_kernel_step and bb.tiles is not None
    bb.boards.copy_(src.boards)
    del src
    bb._mask_valid = False
    idx = torch.arange(0, n, 65537, device=dev)
    raw = np.stack([Oracle.mt_raw(1000 + int(i) % streams, 1024) for i in np_(idx)])
    sample = np_(bb.array[idx])
    acc = torch.zeros(n, dtype=torch.int64, device=dev)
    for t in range(moves):
        bb.apply_action(None)
        acc += bb.step_reward
        a = np_(bb.last_actions[idx])
        res = o.step_batch(sample, a, mode="replay", raw=raw)
        assert not res["status"].any()
        assert np.array_equal(np_(bb.array[idx]), res["boards"])
        assert np.array_equal(np_(bb.step_reward[idx]), res["reward"])
        assert np.array_equal(np_(bb.cascades[idx]), res["cascades"])
        sample = res["boards"]
        st = bb.status
        assert not bool((st & ~(E.ST_NO_LEGAL | E.ST_SHUFFLE_CAP)).any())
        obs = bb.observe(torch.uint8)
        assert int(obs.min()) >= 1 and int(obs.max()) <= 32
        assert int(bb.stream_pos.max()) < 1024 and float(bb.stream_pos.float().mean()) > 4  # words drawn per step
    assert torch.equal(acc, bb.reward)
    assert 1.6 < float(bb.cascades.float().mean()) < 2.0  # the reference's heavier cascade load (Philox: 1.53)


def test_expand_and_greedy_vs_oracle(E):
    """All legal children of every board in one kernel (Node.expand / greedy_action, boardv2.py:209-218)."""
    shape = (9, 9, 6)
    o = Oracle(*shape)
    n = 257
    seeds = np.arange(100, 100 + n, dtype=np.int64)
    bb = E.BatchedBoards(cfg_of(E, shape), n, 5, refill="replay", seeds=seeds, stream_len=2048)
    boards = np_(bb.array)
    child, parent, action = bb.expand()
    parent, action = np_(parent), np_(action)
    legal = o.legal_mask_batch(boards)
    pi, ai = np.nonzero(legal)
    assert np.array_equal(parent, pi) and np.array_equal(action, ai)
    raw = np.stack([Oracle.mt_raw(int(s), 2048) for s in seeds])
    res = o.step_batch(boards[parent], action, mode="replay", raw=raw[parent])
    assert np.array_equal(np_(child.array), res["boards"])
    assert np.array_equal(np_(child.step_reward), res["reward"])
    assert np.array_equal(np_(child.legal_mask()), res["legal"])
    assert (np_(child.moves_left) == 4).all() and np.array_equal(np_(child.reward), res["reward"])
    # greedy: first maximum over ascending legal actions
    want = np.full(n, -1)
    for b in range(n):
        sel = parent == b
        if sel.any():
            r = res["reward"][sel]
            want[b] = action[sel][np.argmax(r)]
    assert np.array_equal(np_(bb.greedy_action()), want)
    # philox mode: children share the parent's (board, step) substream
    pb = E.BatchedBoards(cfg_of(E, shape), n, 5, arrays=boards, key=KEY, board0=50)
    pb.step_ctr = 3
    ch, par, act = pb.expand()
    par, act = np_(par), np_(act)
    for j in range(0, len(par), 37):
        r1 = o.step_batch(boards[par[j]:par[j] + 1], act[j:j + 1], mode="philox", key=KEY, board0=50 + int(par[j]), step_ctr=3)
        assert np.array_equal(np_(ch.array[j:j + 1]), r1["boards"]) and int(ch.step_reward[j]) == r1["reward"][0]


def oracle_pick(o, raw, pos, legal_row):
    """np.random.choice(legal_actions) on the replayed stream at word `pos` -> (action, new position)"""
    import ctypes as C
    rng = o.rng_replay(raw)
    rng.pos = int(pos)
    la = np.flatnonzero(legal_row)
    k = o.L.ecgo_rng_below(C.byref(rng), C.c_uint32(len(la)))
    return int(la[k]), int(rng.pos)


@pytest.mark.parametrize("shape", [(9, 9, 6), (6, 6, 4)])
def test_shared_stream_replay_many_boards(E, shape):
    """SURVEY 8d config 2 "variant": n > 1 boards that all replay ONE MT(cfg.seed) stream (stream_stride = 0) -- the
    reference's situation for every state of one BoardConfig (MCTS leaves, greedy children).  Random picks continue
    the stream behind each board's own refill draws.  Checked against the oracle and against the per-board-stream
    layout (stride > 0) fed with n copies of the same seed."""
    n, moves, seed = 777, 6, 4242
    o = Oracle(*shape)
    rng = np.random.default_rng(9)
    # different boards, same seed: boards reached by Philox play from random starts
    start = E.BatchedBoards(cfg_of(E, shape, seed), n, 3, key=KEY)
    for _ in range(3):
        start.apply_action(None)
    boards = np_(start.array)
    raw = Oracle.mt_raw(seed, 4096)
    shared = E.BatchedBoards(cfg_of(E, shape, seed), n, moves, arrays=boards, refill="replay", seeds=[seed])
    strided = E.BatchedBoards(cfg_of(E, shape, seed), n, moves, arrays=boards, refill="replay", seeds=[seed] * n)
    assert shared.stream_stride == 0 and strided.stream_stride == 4096
    assert shared.stream.numel() == 4096 and np.array_equal(np_(shared.stream).view(np.uint32), raw)
    pos = np.zeros(n, dtype=np.int64)
    for t in range(moves):
        legal = o.legal_mask_batch(boards)
        if t % 2 == 0:
            acts = None  # fused pick inside the step kernel
            want = [oracle_pick(o, raw, pos[i], legal[i])[0] for i in range(n)]
        else:
            acts = np_(shared.random_action())  # separate pick kernel (advances stream_pos; the step restarts it)
            want = [oracle_pick(o, raw, pos[i], legal[i])[0] for i in range(n)]
            assert acts.tolist() == want
        shared.apply_action(acts)
        strided.apply_action(acts)
        a = np_(shared.last_actions)
        assert a.tolist() == want and np.array_equal(a, np_(strided.last_actions))
        res = o.step_batch(boards, a, mode="replay", raw=raw)
        assert not res["status"].any() and not np_(shared.status).any()
        for bb in (shared, strided):
            assert np.array_equal(np_(bb.array), res["boards"])
            assert np.array_equal(np_(bb.step_reward), res["reward"])
            assert np.array_equal(np_(bb.cascades), res["cascades"])
            assert np.array_equal(np_(bb.legal_mask()), res["legal"])
        assert np.array_equal(np_(shared.stream_pos), np_(strided.stream_pos))
        pos = np_(shared.stream_pos).astype(np.int64)
        boards = res["boards"]
    # the same boards in ONE rollout kernel
    roll = E.BatchedBoards(cfg_of(E, shape, seed), n, moves, arrays=np_(start.array), refill="replay", seeds=[seed])
    total = roll.rollout()
    assert np.array_equal(np_(total), np_(shared.reward)) and np.array_equal(np_(roll.array), boards)


def test_replay_children_keep_their_parents_streams(E):
    """expand() in replay mode with one stream per board: child j replays the stream of ITS parent (stream_index)
    from the position its own step reached -- checked by stepping the children (random pick + explicit action) and
    the grandchildren against the oracle."""
    shape = (9, 9, 6)
    o = Oracle(*shape)
    n = 41
    seeds = np.arange(500, 500 + n, dtype=np.int64)
    raw = np.stack([Oracle.mt_raw(int(s), 2048) for s in seeds])
    bb = E.BatchedBoards(cfg_of(E, shape), n, 6, refill="replay", seeds=seeds, stream_len=2048)
    boards = np_(bb.array)
    child, parent, action = bb.expand()
    parent, action = np_(parent), np_(action)
    p = len(parent)
    assert p > 8 * n and np.array_equal(np_(child.stream_index), parent)
    res = o.step_batch(boards[parent], action, mode="replay", raw=raw[parent])
    assert np.array_equal(np_(child.array), res["boards"])
    # stream_pos of a child = words its step drew
    pos = np_(child.stream_pos)
    want_pick = [oracle_pick(o, raw[parent[j]], pos[j], res["legal"][j])[0] for j in range(p)]
    c2 = child.clone()
    assert c2.stream_index is child.stream_index
    c2.apply_action(None)  # np.random.choice continuing the parent's stream
    assert np_(c2.last_actions).tolist() == want_pick
    res2 = o.step_batch(res["boards"], np.array(want_pick), mode="replay", raw=raw[parent])
    assert np.array_equal(np_(c2.array), res2["boards"]) and np.array_equal(np_(c2.step_reward), res2["reward"])
    assert np.array_equal(np_(c2.reward), res["reward"] + res2["reward"])
    # grandchildren: stream_index composes
    g, gpar, gact = child.expand()
    gpar, gact = np_(gpar), np_(gact)
    assert np.array_equal(np_(g.stream_index), parent[gpar])
    sel = np.arange(0, len(gpar), 53)
    res3 = o.step_batch(res["boards"][gpar[sel]], gact[sel], mode="replay", raw=raw[parent[gpar[sel]]])
    assert np.array_equal(np_(g.array)[sel], res3["boards"]) and np.array_equal(np_(g.step_reward)[sel], res3["reward"])
    # whole-episode rollout of the children on their parents' streams == stepping them with random picks
    r = child.clone()
    tot = r.rollout()
    s = child.clone()
    for _ in range(5):
        s.apply_action(None)
    assert np.array_equal(np_(tot) + res["reward"], np_(s.reward)) and np.array_equal(np_(r.array), np_(s.array))


def test_unpack_uint8_partial_tiles_stay_in_bounds(E):
    """uint8 observations are written 4 cells per store: the last word of a partial tile must not run past the
    n*R*C cells that exist (81 cells for n = 1 on 9x9: round-1 advisor finding)."""
    import torch
    for shape in ((9, 9, 6), (5, 5, 2), (7, 7, 5), (6, 6, 4)):
        for n in (1, 2, 3, 33, 35):
            o = Oracle(*shape)
            bb = E.BatchedBoards(cfg_of(E, shape), n, key=KEY)
            want = np_(bb.array).astype(np.uint8)
            cells = n * shape[0] * shape[1]
            buf = torch.full((cells + 64,), 0xAB, dtype=torch.uint8, device=bb.device)
            out = buf[:cells].view(n, shape[0], shape[1])
            bb.observe(torch.uint8, out=out)
            assert np.array_equal(np_(out), want)
            assert bool((buf[cells:] == 0xAB).all()), (shape, n)


GUARDED = ("boards", "mask", "moves_left", "score", "step_reward", "cascades", "flags", "status", "last_actions",
           "_scratch", "stream_pos")


def guard(bb, G=256):
    """Re-home every device buffer the step kernels touch between two sentinel bands.  compute-sanitizer is closed on
    this GPU pool, so out-of-bounds WRITES are caught this way: any store past either end of a buffer lands in a band."""
    import torch
    bands = []
    for name in GUARDED:
        t = getattr(bb, name, None)
        if t is None:
            continue
        fill = 0x5A if t.dtype == torch.uint8 else 0x5A5A5A5A if t.dtype == torch.int32 else 0x5A5A
        big = torch.full((t.numel() + 2 * G,), fill, dtype=t.dtype, device=t.device)
        big[G:G + t.numel()].copy_(t.reshape(-1))
        setattr(bb, name, big[G:G + t.numel()].view(t.shape))
        bands.append((name, big, G, fill))
    return bands


def check_guard(bands):
    for name, big, G, fill in bands:
        assert bool((big[:G] == fill).all()) and bool((big[-G:] == fill).all()), f"out-of-bounds write around {name}"


@pytest.mark.parametrize("shape", [(9, 9, 6), (16, 16, 8), (6, 6, 4), (13, 13, 9)])
@pytest.mark.parametrize("n", [1, 33, 2048 + 13])
def test_no_out_of_bounds_writes(E, shape, n):
    """Every kernel of the step path with its buffers between sentinel bands: two-kernel Philox and replay steps
    (hand-off list), explicit actions, whole-episode rollouts, expand() through src_index, observations."""
    import torch
    cfg = cfg_of(E, shape)
    o = Oracle(*shape)
    for mode in ("philox", "replay"):
        kw = dict(key=KEY) if mode == "philox" else dict(refill="replay", seeds=list(range(7, 7 + n)), stream_len=1024)
        bb = E.BatchedBoards(cfg, n, 5, **kw)
        bands = guard(bb)
        before = np_(bb.array)
        bb.apply_action(None)
        a = np_(bb.last_actions)
        bb.apply_action(bb.random_action())
        child, parent, action = bb.expand()
        cb = guard(child)
        child.apply_action(None)
        roll = bb.clone()
        rb = guard(roll)
        roll.rollout()
        obs = torch.full((n * shape[0] * shape[1] + 512,), 0x5A, dtype=torch.uint8, device=bb.device)
        bb.observe(torch.uint8, out=obs[256:256 + n * shape[0] * shape[1]].view(n, shape[0], shape[1]))
        torch.cuda.synchronize()
        check_guard(bands), check_guard(cb), check_guard(rb)
        assert bool((obs[:256] == 0x5A).all()) and bool((obs[-256:] == 0x5A).all())
        if mode == "philox":  # and the results are still the oracle's
            res = o.step_batch(before, a, mode="philox", key=KEY, board0=0, step_ctr=0)
            again = E.BatchedBoards(cfg, n, 5, arrays=before, key=KEY)
            again.apply_action(a.astype(np.int32))
            assert np.array_equal(np_(again.array), res["boards"])


def test_onehot_observation(E):
    import torch
    for shape in ((9, 9, 6), (16, 16, 8), (6, 6, 4)):
        o = Oracle(*shape)
        rng = np.random.default_rng(3)
        vals = np.array(list(range(0, shape[2] + 1)) + [o.cfg.h_line, o.cfg.v_line, o.cfg.bomb, o.cfg.mega_token])
        b = vals[rng.integers(len(vals), size=(130, shape[0], shape[1]))].astype(np.int64)
        bb = E.BatchedBoards(cfg_of(E, shape), len(b), arrays=b, key=KEY)
        import math
        ch = 2 ** (int(math.ceil(math.log2(shape[2]))) + 2)  # elementCrush.py:66
        want = (b[..., None] == np.arange(ch)[None, None, None, :])  # jax.nn.one_hot: out-of-range -> zeros
        for dt in (torch.float32, torch.uint8, torch.bfloat16, torch.float16):
            got = bb.observe_onehot(dtype=dt)
            assert got.shape == (len(b), shape[0], shape[1], ch)
            assert np.array_equal(got.float().cpu().numpy(), want.astype(np.float32)), (shape, dt)

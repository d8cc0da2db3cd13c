"""The kernel core (csrc/ecg_core.cuh) compiled for the host (tests/hostsim, test-only) against the CPU
oracle and the reference-generated golden vectors.  This is the same __host__ __device__ code the sm_100a
kernels are built from, so the bit-sliced algorithm is checked here without a GPU; the -m gpu tests then
check the CUDA build of it through the C-ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ALL_SHAPES, GOLDEN, SHAPES  # noqa: E402
from hostsim import hostsim as _hostsim  # noqa: E402
from hostsim.hostsim import HostSim  # noqa: E402
from oracle.oracle import Oracle, ST_CASCADE_CAP, ST_SHUFFLE_CAP, ST_STREAM_OVERFLOW, ST_TERMINAL, ST_BAD_ACTION  # noqa: E402

KEY = 0x1234567890ABCDEF


@pytest.fixture(scope="module", autouse=True)
def _host_builds():
    _hostsim.build()  # one library per board size; a fresh checkout compiles them in parallel (about a minute)


def load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_golden_single_steps_replay(shape):
    d = load("steps_%dx%dx%d.npz" % shape)
    h = HostSim(*shape)
    raw = np.stack([Oracle.mt_raw(int(s), 4096) for s in d["seeds"]])
    res = h.step(d["before"].astype(np.int64), d["actions"], mode="replay", raw=raw)
    assert not res["status"].any()
    assert np.array_equal(res["boards"], d["after"])
    assert np.array_equal(res["reward"], d["rewards"])
    assert np.array_equal(res["cascades"], d["cascades"])


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_golden_functions(shape):
    d = load("funcs_%dx%dx%d.npz" % shape)
    o, h = Oracle(*shape), HostSim(*shape)
    b = d["boards"].astype(np.int64)
    legal = np.unpackbits(d["legal"], axis=-1)[..., :o.A].astype(bool)
    assert np.array_equal(h.legal(b), legal)
    mm = np.unpackbits(d["match_mask"], axis=-1)[..., :shape[1]].astype(bool)
    tb = b & o.cfg.type_mask
    mask, spawn = h.matches(tb)
    assert np.array_equal(mask, mm)
    want = d["spawn"].astype(np.int32)
    if o.cfg.type_mask == 15:  # np.clip(next_state, 0, 32) (boardv2.py:163) is folded into the spawn kind
        want = np.minimum(want, 32)
    assert np.array_equal(spawn, want)


def test_golden_shuffle_cases():
    d = load("shuffle.npz")
    for i in range(len(d["actions"])):
        R, Cc, T = (int(x) for x in d["shape"][i])
        h = HostSim(R, Cc, T)
        raw = Oracle.mt_raw(int(d["seeds"][i]), 4096)
        res = h.step(d["before"][i:i + 1, :R, :Cc].astype(np.int64), d["actions"][i:i + 1], mode="replay", raw=raw)
        assert res["status"][0] == 0
        assert np.array_equal(res["boards"][0], d["after"][i, :R, :Cc])
        assert res["reward"][0] == d["rewards"][i] and res["cascades"][0] == d["cascades"][i]


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_golden_episodes_replay(shape):
    """Whole random_task episodes (init board from the MT stream, masked-rejection picks, steps)."""
    d = load("episodes_%dx%dx%d.npz" % shape)
    h = HostSim(*shape)
    seeds = d["seeds"]
    raw = np.stack([Oracle.mt_raw(int(s), 8192) for s in seeds])
    boards, st = h.init(mode="replay", n=len(seeds), raw=raw)
    assert not st.any()
    assert np.array_equal(boards, d["init"])
    for t in range(int(d["moves"])):
        res = h.step(boards, d["actions"][:, t], mode="replay", raw=raw)
        boards = res["boards"]
        assert np.array_equal(boards, d["boards"][:, t])
        assert np.array_equal(res["reward"], d["rewards"][:, t])
        assert np.array_equal(res["cascades"], d["cascades"][:, t])
        if t + 1 < int(d["moves"]):
            legal = np.unpackbits(d["legal"][:, t + 1], axis=-1)[..., :h.A].astype(bool)
            assert np.array_equal(res["legal"], legal)


def _fuzz_boards(rng, o, n):
    R, Cc, T = o.rows, o.cols, o.types
    sp = [o.cfg.h_line, o.cfg.v_line, o.cfg.bomb, o.cfg.mega_token]
    tl = rng.integers(2, T + 1, size=n)
    b = np.stack([rng.integers(1, t + 1, size=(R, Cc)) for t in tl]).astype(np.int64)
    for i in range(0, n, 3):
        for _ in range(int(rng.integers(0, 4))):
            b[i, rng.integers(R), rng.integers(Cc)] = sp[rng.integers(4)]
        if i % 2 == 0:
            b[i, rng.integers(R), rng.integers(Cc)] = 0
    return b, sp


@pytest.mark.parametrize("shape", [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8), (6, 6, 3), (5, 5, 2), (7, 7, 5),
                                   (9, 9, 3), (9, 9, 8), (6, 6, 11)])
def test_fuzz_dense_boards_philox(shape):
    """Dense random boards (few types -> intersecting runs, merged groups, long cascades), specials on
    and off the swapped pair, legal and illegal actions; Philox refill on both sides."""
    rng = np.random.default_rng(shape[0] * 100 + shape[2])
    o, h = Oracle(*shape), HostSim(*shape)
    n = 1500
    b, sp = _fuzz_boards(rng, o, n)
    lo = o.legal_mask_batch(b)
    assert np.array_equal(h.legal(b), lo)
    tb = np.where(b > o.cfg.type_mask, 0, b & o.cfg.type_mask)
    mh, sh = h.matches(tb)
    for i in range(300):
        mo, so, _ = o.matches_and_spawn(tb[i])
        if o.cfg.type_mask == 15:
            so = np.minimum(so, 32)
        assert np.array_equal(mo, mh[i]) and np.array_equal(so, sh[i]), tb[i]
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
    ro = o.step_batch(b, acts, mode="philox", key=KEY, board0=77, step_ctr=5)
    rh = h.step(b, acts, mode="philox", key=KEY, board0=77, step_ctr=5)
    for k in ("boards", "reward", "cascades", "status", "legal"):
        assert np.array_equal(ro[k], rh[k]), k
    # the two-kernel step: common-case build (find_matches<SH, true>), exact build for the boards it hands off
    r2 = h.step(b, acts, mode="philox", key=KEY, board0=77, step_ctr=5, two_pass=True)
    for k in ("boards", "reward", "cascades", "status", "legal"):
        assert np.array_equal(ro[k], r2[k]), k
    assert 0 < r2["handoffs"] < n  # both builds were exercised


@pytest.mark.parametrize("shape", [(9, 9, 6), (9, 9, 8), (6, 6, 4), (12, 12, 7)])
def test_pooled_decomposition_philox(shape):
    """The pooled step kernel's cut of the common-case pass (csrc/ecg_shape_kernels.cu, -DECG_POOL=1): BEGIN, one
    step_iter<DEFER_LEGAL> per pass, the legal swaps at FINISH -- same results as the oracle on natural play and on
    dense boards with planted specials, and cascade_class agrees with the pending match of every pass."""
    o, h = Oracle(*shape), HostSim(*shape)
    n = 1200
    b, _ = h.init(mode="philox", n=n, key=KEY)
    rng = np.random.default_rng(11)
    handoffs = 0
    for step in range(10):
        lo = o.legal_mask_batch(b)
        acts = np.array([rng.choice(np.flatnonzero(m)) if m.any() else 0 for m in lo], dtype=np.int32)
        if step == 5:  # plant specials and an empty cell: the class "everything else", special pairs, extra holes
            sp = [o.cfg.h_line, o.cfg.v_line, o.cfg.bomb, o.cfg.mega_token]
            for i in range(0, n, 3):
                b[i, rng.integers(shape[0]), rng.integers(shape[1])] = sp[rng.integers(4)]
            for i in range(1, n, 7):
                b[i, rng.integers(shape[0]), rng.integers(shape[1])] = 0
        ro = o.step_batch(b, acts, mode="philox", key=KEY, board0=0, step_ctr=step)
        rp = h.step(b, acts, mode="philox", key=KEY, board0=0, step_ctr=step, pooled=True)
        for k in ("boards", "reward", "cascades", "status", "legal"):
            assert np.array_equal(ro[k], rp[k]), (k, step)
        assert rp["class_errors"] == 0
        handoffs += rp["handoffs"]
        b = ro["boards"]
    assert 0 < handoffs < 0.3 * n * 10


@pytest.mark.parametrize("shape", [(9, 9, 6), (6, 6, 4), (12, 12, 7)])
def test_two_pass_episodes_philox(shape):
    """Natural play (random legal actions from fresh boards, 12 steps): most single crossings are plain triples,
    which the common-case build resolves in closed form instead of handing them off."""
    o, h = Oracle(*shape), HostSim(*shape)
    n = 1500
    b, _ = h.init(mode="philox", n=n, key=KEY)
    rng = np.random.default_rng(7)
    handoffs = 0
    for step in range(12):
        lo = o.legal_mask_batch(b)
        acts = np.array([rng.choice(np.flatnonzero(m)) if m.any() else 0 for m in lo], dtype=np.int32)
        ro = o.step_batch(b, acts, mode="philox", key=KEY, board0=0, step_ctr=step)
        r2 = h.step(b, acts, mode="philox", key=KEY, board0=0, step_ctr=step, two_pass=True)
        for k in ("boards", "reward", "cascades", "status", "legal"):
            assert np.array_equal(ro[k], r2[k]), (k, step)
        handoffs += r2["handoffs"]
        b = ro["boards"]
    print("hand-off rate %dx%dx%d: %.2f %%" % (*shape, 100.0 * handoffs / (n * 12)))
    assert 0 < handoffs < (0.04 if shape == (9, 9, 6) else 0.25) * n * 12


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_replay_two_pass_tile_tables(shape):
    """The replay two-kernel step on the host: the common-case pass takes its refill tiles from the per-stream
    tile table (build_replay_tiles / ReplayTileRng) instead of rejection-sampling raw words.  Boards, rewards, cascade
    counts, legal sets AND np.random's position behind the step must equal the exact pass and the reference-generated
    episodes; short streams must overflow exactly like the exact build."""
    d = load("episodes_%dx%dx%d.npz" % shape)
    h = HostSim(*shape)
    seeds = d["seeds"]
    raw = np.stack([Oracle.mt_raw(int(s), 2048) for s in seeds])
    boards = d["init"].astype(np.int64)
    handoffs = 0
    for t in range(int(d["moves"])):
        one = h.step(boards, d["actions"][:, t], mode="replay", raw=raw)
        two = h.step(boards, d["actions"][:, t], mode="replay", raw=raw, two_pass=True)
        handoffs += two["handoffs"]
        for k in ("boards", "reward", "cascades", "status", "legal", "words"):
            assert np.array_equal(one[k], two[k]), (k, t)
        assert np.array_equal(two["boards"], d["boards"][:, t]) and np.array_equal(two["reward"], d["rewards"][:, t])
        boards = two["boards"]
    assert handoffs < len(seeds) * int(d["moves"])  # the tile path was exercised
    # dense boards with planted specials, one shared stream, and a stream that is too short for some of the steps
    rng = np.random.default_rng(shape[0] * 7 + shape[2])
    o = Oracle(*shape)
    b, _ = _fuzz_boards(rng, o, 1200)
    acts = rng.integers(0, o.A, size=len(b))
    for raw1 in (Oracle.mt_raw(99, 2048), Oracle.mt_raw(99, 6), Oracle.mt_raw(5, 1)):
        one = h.step(b, acts, mode="replay", raw=raw1)
        two = h.step(b, acts, mode="replay", raw=raw1, two_pass=True)
        ro = o.step_batch(b, acts, mode="replay", raw=raw1)
        for k in ("boards", "reward", "cascades", "status", "legal"):
            assert np.array_equal(one[k], two[k]) and np.array_equal(ro[k], two[k]), (k, len(raw1))
        ok = (two["status"] & ST_STREAM_OVERFLOW) == 0
        assert np.array_equal(one["words"][ok], two["words"][ok])
        if len(raw1) < 10:
            assert (two["status"] & ST_STREAM_OVERFLOW).any()


def test_caps_and_flags():
    # cascade cap: 9x9 with 2 types never settles
    o, h = Oracle(9, 9, 2), HostSim(9, 9, 2)
    rng = np.random.default_rng(5)
    b = rng.integers(1, 3, size=(8, 9, 9)).astype(np.int64)
    acts = rng.integers(0, o.A, size=8)
    ro = o.step_batch(b, acts, mode="philox", key=KEY)
    rh = h.step(b, acts, mode="philox", key=KEY)
    assert (ro["status"] & ST_CASCADE_CAP).all()
    for k in ("boards", "reward", "cascades", "status", "legal"):
        assert np.array_equal(ro[k], rh[k]), k
    # shuffle cap: the board that hangs the reference (SURVEY 8a ledger)
    o, h = Oracle(9, 9, 6), HostSim(9, 9, 6)
    arr = np.fromfunction(lambda r, c: ((c + 2 * (r % 3)) % 6) + 1, (9, 9), dtype=np.int64).astype(np.int64)[None]
    raw = Oracle.mt_raw(7, 4096)
    ro = o.step_batch(arr, [7], mode="replay", raw=raw)
    rh = h.step(arr, [7], mode="replay", raw=raw)
    assert ro["status"][0] & ST_SHUFFLE_CAP
    for k in ("boards", "reward", "cascades", "status", "legal"):
        assert np.array_equal(ro[k], rh[k]), k
    # terminal boards and bad actions are returned unchanged with a flag
    b = rng.integers(1, 7, size=(4, 9, 9)).astype(np.int64)
    rh = h.step(b, [3, 3, -1, 144], mode="philox", key=KEY, moves_left=[0, 1, 5, 5])
    assert rh["status"][0] == ST_TERMINAL and np.array_equal(rh["boards"][0], b[0])
    assert rh["status"][2] == ST_BAD_ACTION and rh["status"][3] == ST_BAD_ACTION
    assert np.array_equal(rh["boards"][2:], b[2:]) and not rh["reward"][[0, 2, 3]].any()
    assert np.array_equal(rh["legal"][[0, 2, 3]], o.legal_mask_batch(b[[0, 2, 3]]))
    # replay stream too short -> overflow flag on both sides
    raw = Oracle.mt_raw(3, 2)
    b = rng.integers(1, 3, size=(2, 9, 9)).astype(np.int64)
    ro = o.step_batch(b, [0, 1], mode="replay", raw=raw)
    rh = h.step(b, [0, 1], mode="replay", raw=raw)
    assert (ro["status"] & ST_STREAM_OVERFLOW).all() and (rh["status"] & ST_STREAM_OVERFLOW).all()


def test_philox_init_boards_have_no_matches_and_match_oracle_rule():
    h, o = HostSim(9, 9, 6), Oracle(9, 9, 6)
    boards, st = h.init(mode="philox", n=256, key=KEY, board0=1000)
    assert not st.any()
    assert boards.min() >= 1 and boards.max() <= 6
    for b in boards[:64]:
        mask, _, n = o.matches_and_spawn(b)
        assert n == 0
    # same rule as BoardV2.__init__ on the same substream: oracle with the same Philox source
    for i in range(16):
        rng = o.rng_philox(KEY, 1000 + i, 0xFFFFFFFF)
        assert np.array_equal(o.init_board(rng), boards[i])

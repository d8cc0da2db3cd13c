"""The CPU oracle (oracle/ecg_oracle.c) pinned against vectors produced by the UNMODIFIED reference
(scripts/gen_golden.py imports /root/reference; the .npz files are committed)."""
import os

import numpy as np
import pytest

from conftest import ALL_SHAPES, GOLDEN, SHAPES
from oracle.oracle import Oracle, ST_SHUFFLE_CAP


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def unpack_legal(packed, A):
    return np.unpackbits(packed, axis=-1)[..., :A].astype(bool)


def test_rng_matches_numpy_legacy_vectors():
    d = load("rng.npz")
    o = Oracle()
    for i, s in enumerate(d["seeds"]):
        raw = Oracle.mt_raw(int(s), 700)
        assert np.array_equal(raw, d["raw"][i])
        for types in (6, 4, 7, 8):
            r = o.rng_mt(int(s))
            got = [1 + o.L.ecgo_rng_below(r, types) for _ in range(200)] if False else None
        for types in (6, 4, 7, 8):
            r = o.rng_replay(raw)
            import ctypes as C
            got = np.array([1 + o.L.ecgo_rng_below(C.byref(r), C.c_uint32(types)) for _ in range(200)])
            assert np.array_equal(got, d[f"randint{types}"][i]), (s, types)
        import ctypes as C
        r = o.rng_mt(int(s))
        got = [int(o.L.ecgo_rng_below(C.byref(r), C.c_uint32(int(n)))) for n in d["choice_n"]]
        assert got == d["choice"][i].tolist()


def test_rng_row_shuffle_matches_numpy():
    d = load("rng.npz")
    o = Oracle(9, 9, 8)  # type_mask 15: row labels 1..9 are plain tokens, not specials
    for i, s in enumerate(d["seeds"]):
        arr = np.repeat(np.arange(1, 10)[:, None], 9, axis=1).astype(np.int64)  # row r holds r+1
        r = o.rng_mt(int(s))
        out = o.shuffle(r, arr)
        assert np.array_equal(out[:, 0] - 1, d["shuffle9"][i])


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_episodes(shape):
    """random_task trajectories: init board, every chosen action, board, reward, cascade count, legal set."""
    d = load("episodes_%dx%dx%d.npz" % shape)
    o = Oracle(*shape)
    moves = int(d["moves"])
    for e, seed in enumerate(d["seeds"]):
        rng = o.rng_mt(int(seed))
        board = o.init_board(rng)
        assert np.array_equal(board, d["init"][e])
        o.L.ecgo_rng_reseed(__import__("ctypes").byref(rng))
        legal_g = unpack_legal(d["legal"][e], o.A)
        for t in range(moves):
            la = o.legal_actions(board)
            assert la == np.flatnonzero(legal_g[t]).tolist()
            import ctypes as C
            a = la[o.L.ecgo_rng_below(C.byref(rng), C.c_uint32(len(la)))]
            assert a == d["actions"][e, t]
            board, rew, casc, _, st = o.apply_action(rng, board, a)
            assert st == 0
            assert np.array_equal(board, d["boards"][e, t])
            assert rew == d["rewards"][e, t]
            assert casc == d["cascades"][e, t]
        total, steps, fb = o.random_episode(int(seed), moves)
        assert total == int(d["rewards"][e].sum()) and steps == moves
        assert np.array_equal(fb, d["boards"][e, -1])


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_single_steps(shape):
    d = load("steps_%dx%dx%d.npz" % shape)
    o = Oracle(*shape)
    for i in range(len(d["actions"])):
        rng = o.rng_mt(int(d["seeds"][i]))
        nb, rew, casc, draws, st = o.apply_action(rng, d["before"][i].astype(np.int64), int(d["actions"][i]))
        assert st == 0, i
        assert np.array_equal(nb, d["after"][i]), i
        assert rew == d["rewards"][i], i
        assert casc == d["cascades"][i], i
        assert draws == d["ndraws"][i], i
    # batched, threaded entry point gives the same answers
    res = o.step_batch(d["before"].astype(np.int64), d["actions"], mode="mt", raw=d["seeds"])
    assert np.array_equal(res["boards"], d["after"])
    assert np.array_equal(res["reward"], d["rewards"])
    assert np.array_equal(res["cascades"], d["cascades"])


@pytest.mark.parametrize("shape", ALL_SHAPES)
def test_functions(shape):
    d = load("funcs_%dx%dx%d.npz" % shape)
    o = Oracle(*shape)
    R, Cc, _ = shape
    legal_g = unpack_legal(d["legal"], o.A)
    mm_g = np.unpackbits(d["match_mask"], axis=-1)[..., :Cc].astype(bool)
    tm = o.cfg.type_mask
    for i, b in enumerate(d["boards"].astype(np.int64)):
        assert o.legal_actions(b) == np.flatnonzero(legal_g[i]).tolist(), i
        mask, spawn, ng = o.matches_and_spawn(b & tm)
        assert np.array_equal(mask, mm_g[i]), i
        assert np.array_equal(spawn, d["spawn"][i]), i
        assert ng == d["ngroups"][i]
    assert np.array_equal(o.legal_mask_batch(d["boards"].astype(np.int64)), legal_g)


def test_shuffle_cases():
    d = load("shuffle.npz")
    for i in range(len(d["actions"])):
        R, Cc, T = (int(x) for x in d["shape"][i])
        o = Oracle(R, Cc, T)
        rng = o.rng_mt(int(d["seeds"][i]))
        nb, rew, casc, _, st = o.apply_action(rng, d["before"][i, :R, :Cc].astype(np.int64), int(d["actions"][i]))
        assert st == 0
        assert np.array_equal(nb, d["after"][i, :R, :Cc])
        assert rew == d["rewards"][i] and casc == d["cascades"][i]


def test_shuffle_cap_instead_of_hanging():
    """SURVEY 8a ledger: BoardConfig(seed=7), action 7 on ((c + 2*(r%3)) % 6) + 1 spins forever in the
    reference; the oracle (and the engine) stop after ECGO_SHUFFLE_CAP shuffles and flag it."""
    o = Oracle(9, 9, 6)
    arr = np.fromfunction(lambda r, c: ((c + 2 * (r % 3)) % 6) + 1, (9, 9), dtype=np.int64).astype(np.int64)
    rng = o.rng_mt(7)
    nb, rew, casc, _, st = o.apply_action(rng, arr, 7)
    assert st & ST_SHUFFLE_CAP


def test_config_and_codec():
    for (R, Cc, T), (tm, h, v, b, m, A) in {
        (6, 6, 4): (7, 8, 16, 24, 32, 60), (9, 9, 6): (7, 8, 16, 24, 32, 144),
        (12, 12, 7): (7, 8, 16, 24, 32, 264), (16, 16, 8): (15, 16, 32, 48, 64, 480),
    }.items():  # SURVEY 8a row A1 [probed]
        o = Oracle(R, Cc, T)
        c = o.cfg
        assert (c.type_mask, c.h_line, c.v_line, c.bomb, c.mega_token, c.action_space) == (tm, h, v, b, m, A)
        seen = set()
        for a in range(A):
            t1, t2 = o.decode(a)
            assert o.encode(t1, t2) == a
            assert abs(t1[0] - t2[0]) + abs(t1[1] - t2[1]) == 1
            seen.add((t1, t2))
        assert len(seen) == A

#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only import from
/root/reference).  Run in the authoring container only; the GPU box has no /root/reference,
so the vectors are committed.  Usage:  python scripts/gen_golden.py [--out tests/golden]

Fixtures (all integer, stored compactly):
  episodes_{R}x{C}x{T}.npz   random_task trajectories (samplerTasks.py:9-14 with explicit seeds)
  steps_{R}x{C}x{T}.npz      single apply_action cases: (board, action, seed) -> (board', reward, cascades)
                             incl. forced arbitrary/illegal actions and planted special tokens
  funcs_{R}x{C}x{T}.npz      legal_actions / get_matches / get_match_spawn_mask on random boards
  shuffle.npz                no-legal-move boards that go through boardFunctions.shuffle
  rng.npz                    numpy legacy RandomState vectors (seed/randint/choice/shuffle)
"""
import argparse
import os
import signal
import sys

os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from match3tile.boardConfig import BoardConfig  # noqa: E402
from match3tile.boardFunctions import get_match_spawn_mask, get_matches, legal_actions, shuffle  # noqa: E402
from match3tile.boardv2 import BoardV2  # noqa: E402

SHAPES = [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8), (6, 6, 3), (5, 5, 2)]
# round 2: the other square sizes boardConfig accepts (main.py:168 only builds (height, height) boards); smaller files
EXTRA_SHAPES = [(4, 4, 4), (7, 7, 5), (8, 8, 5), (10, 10, 6), (11, 11, 7), (13, 13, 9), (14, 14, 6), (15, 15, 8)]


class Timeout(Exception):
    pass


def _alarm(signum, frame):
    raise Timeout()


class Recorder:
    """Counts cascade iterations (np.clip is called once per iteration, boardv2.py:163) and
    records refill draws (np.random.randint inside apply_action, boardv2.py:172)."""

    def __enter__(self):
        self.clips = 0
        self.draws = []
        self._clip, self._randint = np.clip, np.random.randint

        def clip(*a, **k):
            self.clips += 1
            return self._clip(*a, **k)

        def randint(*a, **k):
            out = self._randint(*a, **k)
            self.draws.extend(np.asarray(out).ravel().tolist())
            return out

        np.clip, np.random.randint = clip, randint
        return self

    def __exit__(self, *exc):
        np.clip, np.random.randint = self._clip, self._randint


def apply(cfg, arr, action, timeout=2):
    """reference apply_action on an explicit array -> (next, step_reward, cascades, draws)"""
    st = BoardV2(5, cfg, np.array(arr, dtype=np.int64))
    signal.signal(signal.SIGALRM, _alarm)
    signal.alarm(timeout)
    try:
        with Recorder() as rec:
            nxt = st.apply_action(int(action))
    finally:
        signal.alarm(0)
    return nxt.array.copy(), int(nxt.reward), rec.clips, list(rec.draws)


def gen_episodes(shape, seeds, moves=20):
    R, Cc, T = shape
    init, actions, boards, rewards, casc, legal = [], [], [], [], [], []
    kept = []
    for s in seeds:
        cfg = BoardConfig(seed=int(s), rows=R, columns=Cc, types=T)
        signal.signal(signal.SIGALRM, _alarm)
        signal.alarm(20)  # tiny boards can spin in the reference's shuffle loop (boardv2.py:188-194): skip that seed
        try:
            state = BoardV2(moves, cfg)
            i0 = state.array.copy()
            np.random.seed(cfg.seed)
            ea, eb, er, ec, el = [], [], [], [], []
            while not state.is_terminal:
                la = list(state.legal_actions)
                m = np.zeros(cfg.action_space, dtype=np.uint8)
                m[la] = 1
                a = np.random.choice(la)
                prev = state.reward
                with Recorder() as rec:
                    state = state.apply_action(a)
                ea.append(int(a)); eb.append(state.array.copy()); er.append(int(state.reward - prev))
                ec.append(rec.clips); el.append(m)
        except (Timeout, ValueError):  # ValueError: np.random.choice on an empty legal set
            continue
        finally:
            signal.alarm(0)
        kept.append(s); init.append(i0)
        actions.append(ea); boards.append(eb); rewards.append(er); casc.append(ec); legal.append(el)
    seeds = kept
    return dict(seeds=np.array(seeds, dtype=np.uint32), init=np.array(init, dtype=np.int8),
                actions=np.array(actions, dtype=np.int16), boards=np.array(boards, dtype=np.int8),
                rewards=np.array(rewards, dtype=np.int32), cascades=np.array(casc, dtype=np.int16),
                legal=np.packbits(np.array(legal, dtype=np.uint8), axis=-1), moves=np.int32(moves))


def plant_specials(rng, cfg, arr, action, mode):
    """Put special tokens on the swapped cells and/or elsewhere."""
    specials = [cfg.h_line, cfg.v_line, cfg.bomb, cfg.mega_token]
    (r1, c1), (r2, c2) = cfg.actions[action]
    arr = arr.copy()
    if mode in (0, 1):
        arr[r1, c1] = specials[rng.integers(4)]
    if mode in (0, 2):
        arr[r2, c2] = specials[rng.integers(4)]
    for _ in range(int(rng.integers(0, 4))):
        arr[rng.integers(cfg.rows), rng.integers(cfg.columns)] = specials[rng.integers(4)]
    if rng.integers(5) == 0:
        arr[rng.integers(cfg.rows), rng.integers(cfg.columns)] = 0
    return arr


def gen_steps(shape, n_random, n_special, seed0):
    R, Cc, T = shape
    rng = np.random.default_rng(seed0)
    before, acts, seeds, after, rew, casc, ndraws = [], [], [], [], [], [], []

    def record(cfg, arr, a):
        try:
            nxt, r, c, d = apply(cfg, arr, a)
        except Timeout:
            return
        before.append(arr.copy()); acts.append(a); seeds.append(cfg.seed); after.append(nxt)
        rew.append(r); casc.append(c); ndraws.append(len(d))

    # (a) boards reached by play, 1 in 4 actions arbitrary (usually illegal)
    ep = 0
    while len(before) < n_random:
        ep += 1
        cfg = BoardConfig(seed=int(rng.integers(1, 2**31 - 1)), rows=R, columns=Cc, types=T)
        state = BoardV2(12, cfg)
        while not state.is_terminal and len(before) < n_random:
            la = state.legal_actions
            a = int(rng.integers(cfg.action_space)) if (rng.integers(4) == 0 or not la) else int(la[rng.integers(len(la))])
            record(cfg, state.array, a)
            signal.signal(signal.SIGALRM, _alarm)
            signal.alarm(5)  # small boards can spin in the reference's shuffle loop: drop the episode
            try:
                state = state.apply_action(a)
            except Timeout:
                break
            finally:
                signal.alarm(0)
    # (b) planted specials (all pair branches boardv2.py:81-136 + trigger pass :141-154)
    k = 0
    while k < n_special:
        cfg = BoardConfig(seed=int(rng.integers(1, 2**31 - 1)), rows=R, columns=Cc, types=T)
        arr = BoardV2(5, cfg).array
        a = int(rng.integers(cfg.action_space))
        arr = plant_specials(rng, cfg, arr, a, int(rng.integers(4)))
        n0 = len(before)
        record(cfg, arr, a)
        k += len(before) - n0
    return dict(before=np.array(before, dtype=np.int8), actions=np.array(acts, dtype=np.int16),
                seeds=np.array(seeds, dtype=np.uint32), after=np.array(after, dtype=np.int8),
                rewards=np.array(rew, dtype=np.int32), cascades=np.array(casc, dtype=np.int16),
                ndraws=np.array(ndraws, dtype=np.int16))


def gen_funcs(shape, n, seed0):
    R, Cc, T = shape
    rng = np.random.default_rng(seed0)
    cfg = BoardConfig(seed=1, rows=R, columns=Cc, types=T)
    specials = [cfg.h_line, cfg.v_line, cfg.bomb, cfg.mega_token]
    boards, legal, mmask, spawn, ngroups = [], [], [], [], []
    for i in range(n):
        t = T if i % 3 else max(2, T // 2)  # fewer types -> dense, intersecting matches
        arr = rng.integers(1, t + 1, size=(R, Cc)).astype(np.int64)
        for _ in range(int(rng.integers(0, 4))):
            arr[rng.integers(R), rng.integers(Cc)] = specials[rng.integers(4)]
        if i % 5 == 0:
            arr[rng.integers(R), rng.integers(Cc)] = 0
        la = legal_actions(cfg, arr)
        m = np.zeros(cfg.action_space, dtype=np.uint8)
        m[la] = 1
        tb = arr & cfg.type_mask
        zm, matches = get_matches(tb)
        sp = get_match_spawn_mask(cfg, matches)
        boards.append(arr); legal.append(m); mmask.append(zm.astype(np.uint8)); spawn.append(sp)
        ngroups.append(len(matches))
    return dict(boards=np.array(boards, dtype=np.int8), legal=np.packbits(np.array(legal), axis=-1),
                match_mask=np.packbits(np.array(mmask), axis=-1), spawn=np.array(spawn, dtype=np.int8),
                ngroups=np.array(ngroups, dtype=np.int16))


def gen_shuffle(seed0, want=40):
    """Boards with no legal move whose step goes through 1..k shuffles and terminates."""
    rng = np.random.default_rng(seed0)
    out = dict(shape=[], before=[], actions=[], seeds=[], after=[], rewards=[], cascades=[])
    tries = 0
    while len(out["before"]) < want and tries < 40000:
        tries += 1
        R, Cc, T = [(5, 5, 4), (6, 6, 5), (5, 5, 5), (6, 6, 6)][tries % 4]
        cfg = BoardConfig(seed=int(rng.integers(1, 2**31 - 1)), rows=R, columns=Cc, types=T)
        # diagonal-stripe boards have no runs and few/no legal moves; perturb a little
        k = int(rng.integers(1, 4))
        arr = np.fromfunction(lambda r, c: ((c + k * r) % T) + 1, (R, Cc), dtype=np.int64).astype(np.int64)
        if rng.integers(3) == 0:
            perm = rng.permutation(T) + 1
            arr = perm[arr - 1]
        if legal_actions(cfg, arr):
            continue
        a = int(rng.integers(cfg.action_space))
        calls = {"n": 0}
        orig = np.random.shuffle

        def counting(x):
            calls["n"] += 1
            return orig(x)

        np.random.shuffle = counting
        try:
            nxt, r, c, d = apply(cfg, arr, a, timeout=1)
        except Timeout:
            continue
        finally:
            np.random.shuffle = orig
        if calls["n"] == 0 or calls["n"] > 60:
            continue
        pad = np.zeros((6, 6), dtype=np.int8)
        pb, pa = pad.copy(), pad.copy()
        pb[:R, :Cc] = arr; pa[:R, :Cc] = nxt
        out["shape"].append((R, Cc, T)); out["before"].append(pb); out["actions"].append(a)
        out["seeds"].append(cfg.seed); out["after"].append(pa); out["rewards"].append(r); out["cascades"].append(c)
    return {k: np.array(v) for k, v in out.items()}


def gen_rng():
    """numpy legacy RandomState vectors: raw u32 stream, randint, choice, row shuffle."""
    seeds = [1, 2, 7, 12345, 2**31 - 2, 4294967295]
    raw, r6, r4, r7, r8, ch, sh = [], [], [], [], [], [], []
    for s in seeds:
        np.random.seed(s)
        raw.append(np.random.randint(0, 2**32, size=700, dtype=np.uint64).astype(np.uint32))
        for types, dst in ((6, r6), (4, r4), (7, r7), (8, r8)):
            np.random.seed(s)
            dst.append(np.random.randint(1, types + 1, size=200))
        np.random.seed(s)
        ch.append([int(np.random.choice(list(range(n)))) for n in (1, 2, 3, 5, 16, 17, 33, 144, 1, 9)])
        np.random.seed(s)
        a = np.arange(9 * 2).reshape(9, 2)
        np.random.shuffle(a)
        sh.append(a[:, 0] // 2)
    return dict(seeds=np.array(seeds, dtype=np.uint64), raw=np.array(raw), randint6=np.array(r6),
                randint4=np.array(r4), randint7=np.array(r7), randint8=np.array(r8),
                choice_n=np.array([1, 2, 3, 5, 16, 17, 33, 144, 1, 9]), choice=np.array(ch), shuffle9=np.array(sh))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--extra-only", action="store_true", help="only the round-2 EXTRA_SHAPES fixtures")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    sc = args.scale
    if not args.extra_only:
        np.savez_compressed(os.path.join(args.out, "rng.npz"), **gen_rng())
        print("rng done", flush=True)
    for shape in EXTRA_SHAPES:
        tag = "%dx%dx%d" % shape
        np.savez_compressed(os.path.join(args.out, f"episodes_{tag}.npz"), **gen_episodes(shape, list(range(1, int(6 * sc) + 1))))
        seed0 = shape[0] * 10000 + shape[1] * 100 + shape[2]
        np.savez_compressed(os.path.join(args.out, f"steps_{tag}.npz"), **gen_steps(shape, int(100 * sc), int(80 * sc), seed0=seed0))
        np.savez_compressed(os.path.join(args.out, f"funcs_{tag}.npz"), **gen_funcs(shape, int(80 * sc), seed0=seed0 + 1))
        print(tag, "done", flush=True)
    if args.extra_only:
        return
    for shape in SHAPES:
        tag = "%dx%dx%d" % shape
        main_shape = shape == (9, 9, 6)
        n_ep = int((48 if main_shape else 12) * sc)
        np.savez_compressed(os.path.join(args.out, f"episodes_{tag}.npz"), **gen_episodes(shape, list(range(1, n_ep + 1))))
        np.savez_compressed(os.path.join(args.out, f"steps_{tag}.npz"),
                            **gen_steps(shape, int((600 if main_shape else 200) * sc),
                                        int((400 if main_shape else 150) * sc), seed0=shape[0] * 10000 + shape[1] * 100 + shape[2]))
        np.savez_compressed(os.path.join(args.out, f"funcs_{tag}.npz"),
                            **gen_funcs(shape, int((400 if main_shape else 150) * sc), seed0=shape[0] * 10000 + shape[1] * 100 + shape[2] + 1))
        print(tag, "done", flush=True)
    np.savez_compressed(os.path.join(args.out, "shuffle.npz"), **gen_shuffle(20261018))
    print("shuffle done", flush=True)


if __name__ == "__main__":
    main()

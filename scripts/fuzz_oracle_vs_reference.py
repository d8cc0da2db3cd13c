#!/usr/bin/env python
"""Differential fuzz: CPU oracle vs the imported reference (authoring container only).
   python scripts/fuzz_oracle_vs_reference.py [n_steps] [seed]"""
import os
import signal
import sys

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
from match3tile.boardConfig import BoardConfig  # noqa: E402
from match3tile.boardFunctions import get_match_spawn_mask, get_matches, legal_actions  # noqa: E402
from match3tile.boardv2 import BoardV2  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

SHAPES = [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8), (6, 6, 3), (5, 5, 2), (7, 7, 5), (9, 9, 3)]


class Timeout(Exception):
    pass


def _alarm(s, f):
    raise Timeout()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    signal.signal(signal.SIGALRM, _alarm)
    done = bad = skipped = 0
    while done < n:
        R, Cc, T = SHAPES[int(rng.integers(len(SHAPES)))]
        o = Oracle(R, Cc, T)
        cfg = BoardConfig(seed=int(rng.integers(1, 2**32 - 1)), rows=R, columns=Cc, types=T)
        state = BoardV2(15, cfg)
        assert np.array_equal(state.array, o.init_board(o.rng_mt(cfg.seed)))
        specials = [cfg.h_line, cfg.v_line, cfg.bomb, cfg.mega_token]
        while not state.is_terminal:
            arr = state.array.copy()
            mode = int(rng.integers(8))
            if mode == 0:  # plant specials (typeless)
                for _ in range(int(rng.integers(1, 4))):
                    arr[rng.integers(R), rng.integers(Cc)] = specials[rng.integers(4)]
            elif mode == 1:  # typed specials / zeros: outside the engine's domain, inside the oracle's
                for _ in range(int(rng.integers(1, 4))):
                    arr[rng.integers(R), rng.integers(Cc)] = specials[rng.integers(4)] + int(rng.integers(0, T + 1))
                arr[rng.integers(R), rng.integers(Cc)] = 0
            st = BoardV2(state.n_actions, cfg, arr)
            la = st.legal_actions
            assert la == o.legal_actions(arr), (arr, la)
            tb = arr & cfg.type_mask
            zm, m = get_matches(tb)
            mask, spawn, ng = o.matches_and_spawn(tb)
            assert np.array_equal(zm, mask) and np.array_equal(get_match_spawn_mask(cfg, m), spawn) and ng == len(m)
            a = int(rng.integers(cfg.action_space)) if (rng.integers(5) == 0 or not la) else int(la[rng.integers(len(la))])
            if mode in (0, 1) and rng.integers(2):  # put specials on the swapped pair
                (r1, c1), (r2, c2) = cfg.actions[a]
                arr[r1, c1] = specials[rng.integers(4)]
                if rng.integers(2):
                    arr[r2, c2] = specials[rng.integers(4)]
                st = BoardV2(state.n_actions, cfg, arr)
            signal.alarm(2)
            try:
                nxt = st.apply_action(a)
            except Timeout:
                skipped += 1
                break
            finally:
                signal.alarm(0)
            nb, rew, casc, draws, status = o.apply_action(o.rng_mt(cfg.seed), arr, a)
            if not (np.array_equal(nb, nxt.array) and rew == nxt.reward and status == 0):
                bad += 1
                print("MISMATCH", (R, Cc, T), cfg.seed, a, arr.tolist())
            state = BoardV2(nxt.n_actions, cfg, nxt.array)
            done += 1
    print(f"steps={done} mismatches={bad} skipped_hangs={skipped}")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

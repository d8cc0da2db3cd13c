"""GPU-batched rollouts behind the reference's MCTS call shape (mctslib/standard/mcts.py)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.oracle import Oracle  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ecg_b200
    return ecg_b200


def test_replicated_leaf_rollouts_match_oracle(E):
    from importlib import import_module
    mcts = import_module("element-crush-gym_b200.mcts")
    cfg = E.BoardConfig(seed=11)
    st = E.BoardV2(6, cfg)
    st = st.apply_action(st.legal_actions[0])
    n, key, board0 = 4096, 0xABCDEF, 777
    batch = mcts.replicate(st, n, key=key, board0=board0)
    arr = batch.array.cpu().numpy()
    assert (arr == st.array[None]).all()
    total = batch.rollout().cpu().numpy()
    o = Oracle(9, 9, 6)
    _, want, steps = o.philox_episode_batch(np.repeat(st.array[None], n, axis=0), key, board0, st.n_actions)
    assert np.array_equal(total, want)
    assert np.array_equal(batch.rollout_steps.cpu().numpy(), steps)


def test_mcts_call_shape_and_counts(E):
    cfg = E.BoardConfig(seed=3)
    state = E.BoardV2(4, cfg)
    legal = state.legal_actions
    sims, leaves = 6, 2048
    m = E.BatchedRolloutMCTS(state, 2, sims, False, leaves=leaves, key=5)
    action, value, policies = m()
    assert action in legal
    assert isinstance(value, int) and value >= 0
    assert len(policies) >= 1 and abs(sum(policies) - 1.0) < 1e-9  # every simulation passes through one root child
    assert m.env_steps > 0 and m._root.visits % leaves == 0
    # tree reuse: the new root is the chosen child, already visited
    assert m._root.parent is None and m._root.visits >= leaves
    action2, value2, policies2 = m()
    assert action2 in m._root.parent.state.legal_actions if m._root.parent else True


def stub_value(state):
    """scripts/gen_golden_mcts.py: the deterministic stand-in for MCTS.rollout used to record the reference's tree"""
    a = np.asarray(state.array, dtype=np.int64)
    w = np.arange(1, a.size + 1, dtype=np.int64).reshape(a.shape)
    return int(state.reward) + int((a * w).sum() % 1009)


def test_host_tree_matches_reference_golden(E):
    """the host tree (UCB1 with c = n_actions, pop-largest expansion, policies in insertion order,
    c = 0 value descent, tree re-use) against what the UNMODIFIED reference MCTS returned with the same stubbed
    rollout (tests/golden/mcts_tree.json, scripts/gen_golden_mcts.py)."""
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcts_tree.json")))

    class Stub(E.BatchedRolloutMCTS):
        def rollout(self, state):
            return stub_value(state), 1

    assert len(cases) >= 6
    for case in cases:
        cfg = E.BoardConfig(seed=case["seed"], rows=case["rows"], columns=case["rows"], types=case["types"])
        state = E.BoardV2(case["moves"], cfg)
        m = Stub(state, 3.0, case["simulations"], False)
        for want in case["calls"]:
            root = m._root
            action, value, policies = m()
            assert action == want["action"] and value == want["value"], (case["seed"], action, value)
            assert policies == want["policies"]  # same float arithmetic, same order: exact
            assert list(root.children.keys()) == want["child_actions"] == m.last_root_actions
            assert [c.visits for c in root.children.values()] == want["child_visits"]
            assert [c.reward for c in root.children.values()] == want["child_rewards"]
            assert root.visits == want["root_visits"] and root.reward == want["root_reward"]


def reference_rollout(o, seed, arr, first_action, moves):
    """mctslib/standard/mcts.py:16-18 after its first pick, on the oracle: every apply_action reseeds MT(cfg.seed)
    (boardv2.py:46), every later np.random.choice continues that stream behind the step's refill draws."""
    rng = o.rng_mt(seed)
    total, a = 0, first_action
    for t in range(moves):
        arr, r, _, _, st = o.apply_action(rng, arr, a)
        assert st == 0
        total += r
        la = o.legal_actions(arr)
        if t + 1 < moves:
            a = la[o.L.ecgo_rng_below(C.byref(rng), C.c_uint32(len(la)))]
    return total


def test_replay_rollouts_reproduce_reference_dynamics(E):
    """refill="replay": A12 -- the reference's rollout is a deterministic function of its first action.  Checked
    against the oracle for every legal first action, and against the reference-generated random_task episodes
    (whose first pick is the deterministic=True pick: np.random.seed(cfg.seed); np.random.choice(legal))."""
    import torch
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "episodes_9x9x6.npz"))
    o = Oracle(9, 9, 6)
    moves = int(d["moves"])
    for e in (0, 3, 7):
        seed = int(d["seeds"][e])
        cfg = E.BoardConfig(seed=seed)
        state = E.BoardV2(moves, cfg)
        m = E.BatchedRolloutMCTS(state, 3, 1, False, deterministic=True, leaves=8, refill="replay")
        returns = m._first_action_returns(state)[0].cpu().numpy()
        legal = state.legal_actions
        assert len(returns) == len(legal)
        golden_total = int(d["rewards"][e].sum())
        assert returns[legal.index(int(d["actions"][e, 0]))] == golden_total
        for j in range(0, len(legal), 3):
            assert returns[j] == reference_rollout(o, seed, state.array, legal[j], moves)
        rsum, n = m.rollout(state)
        assert (rsum, n) == (8 * golden_total, 8)
        # a deeper state with points already collected: the return includes state.reward
        s2 = state.apply_action(int(d["actions"][e, 0])).apply_action(int(d["actions"][e, 1]))
        r2 = m._first_action_returns(s2)[0].cpu().numpy()
        assert r2[s2.legal_actions.index(int(d["actions"][e, 2]))] == golden_total
        # non-deterministic: first picks uniform over the legal set, sums reproducible from (key, simulation)
        m2 = E.BatchedRolloutMCTS(state, 3, 1, False, leaves=4096, key=9, refill="replay")
        rsum2, n2 = m2.rollout(state)
        assert n2 == 4096 and 4096 * returns.min() <= rsum2 <= 4096 * returns.max()
        assert abs(rsum2 / 4096 - returns.mean()) < 4 * returns.std() / 64 + 1e-9
    action, value, policies = E.BatchedRolloutMCTS(state, 3, 12, False, leaves=64, refill="replay")()
    assert action in state.legal_actions and abs(sum(policies) - 1.0) < 1e-9
    torch.cuda.synchronize()


def test_unmodified_random_task_loop_matches_reference(E):
    """samplerTasks.random_task (:9-14) run UNMODIFIED on the drop-in state -- np.random.choice on the host's
    global generator -- picks the reference's actions: apply_action leaves np.random where the reference leaves it
    (np.random.seed(cfg.seed) + the words the step drew, boardv2.py:46, :172)."""
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "episodes_6x6x4.npz"))
    for e in (1, 5):
        cfg = E.BoardConfig(seed=int(d["seeds"][e]), rows=6, columns=6, types=4)
        state = E.BoardV2(int(d["moves"]), cfg)
        np.random.seed(state.cfg.seed)
        t = 0
        while not state.is_terminal:
            a = np.random.choice(state.legal_actions)
            assert int(a) == int(d["actions"][e, t])
            state = state.apply_action(a)
            assert np.array_equal(state.array, d["boards"][e, t])
            t += 1
        assert state.reward == int(d["rewards"][e].sum())

"""GPU-batched rollouts behind the reference's MCTS call shape (mctslib/standard/mcts.py)."""
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
    assert m.env_steps > 0
    # tree reuse: the new root is the chosen child, already visited
    assert m._root.parent is None and m._root.visits >= leaves
    action2, value2, policies2 = m()
    assert action2 in m._root.parent.state.legal_actions if m._root.parent else True

"""The dataset path (reference dataset.py): mirror / type-switch augmentation and MCTS self-play samples.
Golden vectors: tests/golden/dataset_mirror.npz, produced by the unmodified reference's Dataset.mirror
(scripts/gen_golden_dataset.py)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "dataset_mirror.npz"))
SHAPES = [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8), (5, 5, 2)]


@pytest.fixture(scope="module")
def E():
    import ecg_b200
    return ecg_b200


@pytest.mark.parametrize("shape", SHAPES)
def test_mirror_action_permutation_matches_reference(E, shape):
    R, Cc, T = shape
    cfg = E.BoardConfig(seed=1, rows=R, columns=Cc, types=T)
    perm = E.dataset.mirror_actions(cfg)
    assert np.array_equal(perm, GOLD[f"perm_{R}x{Cc}x{T}"])
    assert np.array_equal(perm[perm], np.arange(cfg.action_space))  # an involution


def test_augment_argument_checks_without_gpu(E):
    N = E._native
    cfg = E.BoardConfig(seed=1)
    L = N.lib()
    assert L.ecg_augment(C.byref(cfg.native), None, None, 1, None, 4, None) != 0
    bad = bytes([1, 1, 2, 3, 4, 5])  # not a permutation
    assert L.ecg_augment(C.byref(cfg.native), C.c_void_p(16), C.c_void_p(16), 0, bad, 0, None) != 0
    assert b"permutation" in L.ecg_last_error()
    ok = bytes([2, 1, 3, 4, 5, 6])
    assert L.ecg_augment(C.byref(cfg.native), C.c_void_p(16), C.c_void_p(16), 0, ok, 0, None) == 0  # n == 0: no launch


@pytest.fixture(scope="module")
def cuda(E):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES)
def test_mirror_boards_and_policies_match_reference(E, cuda, shape):
    torch = cuda
    R, Cc, T = shape
    tag = f"{R}x{Cc}x{T}"
    cfg = E.BoardConfig(seed=1, rows=R, columns=Cc, types=T)
    obs = GOLD[f"obs_{tag}"]
    b = E.BatchedBoards(cfg, obs.shape[0], 20, refill="philox", key=1, arrays=obs)
    m = E.dataset.augment_boards(b, mirror=True)
    assert np.array_equal(m.array.cpu().numpy(), GOLD[f"mobs_{tag}"])
    assert np.array_equal(b.array.cpu().numpy(), obs)  # the source batch is untouched
    pol = torch.as_tensor(GOLD[f"pol_{tag}"]).cuda()
    assert np.array_equal(E.dataset.mirror_policies(cfg, pol).cpu().numpy(), GOLD[f"mpol_{tag}"])
    # in place (boards_out == boards_in) and twice = identity
    N = E._native
    from importlib import import_module
    bd = import_module("element-crush-gym_b200.boards")
    for _ in range(2):
        N.check(b.L.ecg_augment(C.byref(b.nat), bd._ptr(b.boards), bd._ptr(b.boards), 1, None, b.n, None))
    torch.cuda.synchronize()
    assert np.array_equal(b.array.cpu().numpy(), obs)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(9, 9, 6), (16, 16, 8)])
def test_type_switch_permutes_plain_tokens_only(E, cuda, shape):
    R, Cc, T = shape
    tag = f"{R}x{Cc}x{T}"
    cfg = E.BoardConfig(seed=1, rows=R, columns=Cc, types=T)
    obs = GOLD[f"obs_{tag}"]
    rng = np.random.default_rng(5)
    perm = rng.permutation(np.arange(1, T + 1))
    b = E.BatchedBoards(cfg, obs.shape[0], 20, refill="philox", key=1, arrays=obs)
    got = E.dataset.augment_boards(b, mirror=True, type_perm=perm).array.cpu().numpy()
    lut = np.arange(cfg.mega_token + 1)
    lut[1:T + 1] = perm
    assert np.array_equal(got, lut[np.fliplr(obs.transpose(1, 2, 0)).transpose(2, 0, 1)])
    with pytest.raises(Exception):
        E.dataset.augment_boards(b, type_perm=[1] * T)


@pytest.mark.gpu
def test_dataset_sample_and_split(E, cuda, tmp_path):
    torch = cuda
    cfg = E.BoardConfig(seed=9)
    ds = E.Dataset(cfg, moves=4, simulations=6, leaves=512, key=77)
    ds.sample(20, caching=True, directory=str(tmp_path))
    d = ds.dataset
    n = len(d["values"])
    assert n >= 20 and n % 4 == 0 and len(d["observations"]) == len(d["policies"]) == n
    assert os.path.isfile(tmp_path / "(9, 9, 6).ds")  # the reference's cache name (dataset.py:67)
    for k in range(0, n, 4):  # one value per move of an episode: the final reward (dataset.py:41)
        assert len(set(d["values"][k:k + 4])) == 1
    for o, p in zip(d["observations"], d["policies"]):
        assert o.shape == (9, 9) and o.dtype == np.int64 and p.shape == (cfg.action_space,)
        assert 0 < p.sum() <= 1.0 + 1e-9
    # a second Dataset reads the cache instead of sampling again
    ds2 = E.Dataset(cfg, moves=4, simulations=6, leaves=512).sample(20, caching=True, directory=str(tmp_path))
    assert len(ds2.dataset["values"]) == n
    obs, pol, val = ds.with_mirroring(True).with_type_switching(True, 3).tensors()
    assert obs.shape == (20 * 3 * 2, 9, 9) and pol.shape == (120, cfg.action_space) and val.shape == (120,)
    base = np.stack(d["observations"][:20])
    assert np.array_equal(obs[:20].cpu().numpy(), base)
    assert np.array_equal(obs[60:80].cpu().numpy(), base[:, :, ::-1])  # mirror image of the plain samples
    train, test = ds.with_batching(16).get_split(0.8)
    assert sum(len(b["values"]) for b in train) == 96 and sum(len(b["values"]) for b in test) == 24
    assert float(max(b["values"].max() for b in train + test)) <= 1.0

"""CPU-side checks of the boundary: libecg.so loads, exports every symbol include/ecg.h declares, and its
argument checking / config arithmetic work without a GPU (no kernel is launched here)."""
import ctypes as C
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def E():
    import ecg_b200
    return ecg_b200


def test_library_exports_every_declared_symbol(E):
    hdr = open(os.path.join(ROOT, "include", "ecg.h")).read()
    declared = set(re.findall(r"\b(ecg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    L = C.CDLL(E._native.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(E._native.EXPORTS), declared ^ set(E._native.EXPORTS)
    assert E._native.lib().ecg_version() == 106


def test_struct_layouts_match_header(E):
    N = E._native
    assert C.sizeof(N.Config) == 14 * 4
    assert C.sizeof(N.Refill) == 80 and N.Refill.stream.offset == 8 and N.Refill.philox_key.offset == 32
    assert N.Refill.stream_index.offset == 56 and N.Refill.tiles.offset == 64 and N.Refill.tile_wpos.offset == 72
    assert C.sizeof(N.StepIO) == 12 * 8 + 8 + 8 + 8 and N.StepIO.env_goal.offset == 96 and N.StepIO.src_index.offset == 104
    L = N.lib()
    for which, st in enumerate((N.Config, N.Refill, N.StepIO)):  # ECG_SIZEOF_CONFIG / _REFILL / _STEP_IO
        assert L.ecg_sizeof(which) == C.sizeof(st)
    assert L.ecg_sizeof(3) == -1


def test_integration_md_stubs_match_the_library(E):
    """The ctypes stubs INTEGRATION.md tells a maintainer to copy must be the structs the library reads: a short
    ecg_step_io makes ecg_step read past the caller's struct (round-1 finding: 104 B documented, 120 B read)."""
    import re
    N = E._native
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = next(b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "class EcgStepIO" in b)
    stubs_src = block.split("\nL = C.CDLL")[0]  # the import line and the three class statements
    ns = {}
    exec(compile(stubs_src.replace(", torch", ""), "INTEGRATION.md", "exec"), ns)
    L = N.lib()
    for which, (name, mine) in enumerate((("EcgConfig", N.Config), ("EcgRefill", N.Refill), ("EcgStepIO", N.StepIO))):
        stub = ns[name]
        assert C.sizeof(stub) == C.sizeof(mine) == L.ecg_sizeof(which), name
        assert [(f[0], getattr(stub, f[0]).offset, getattr(stub, f[0]).size) for f in stub._fields_] == \
               [(f[0], getattr(mine, f[0]).offset, getattr(mine, f[0]).size) for f in mine._fields_], name


def test_config_matches_reference_constants(E):
    # SURVEY 8a row A1 [probed on the reference]
    for (R, T), want in {(6, 4): (7, 8, 16, 24, 32, 60), (9, 6): (7, 8, 16, 24, 32, 144),
                         (12, 7): (7, 8, 16, 24, 32, 264), (16, 8): (15, 16, 32, 48, 64, 480)}.items():
        c = E.BoardConfig(seed=1, rows=R, columns=R, types=T)
        assert (c.type_mask, c.h_line, c.v_line, c.bomb, c.mega_token, c.action_space) == want
        assert c.shape == (R, R) and len(c.actions) == c.action_space
        for a, (t1, t2) in c.actions.items():
            assert c.encode(t1, t2) == a == c.encode(t2, t1)
    assert E.BoardConfig(seed=0).seed != 0 and E.BoardConfig().seed  # boardConfig.py:34
    c = E.BoardConfig(seed=1)
    assert c.native.board_words == 12 and c.native.mask_words == 6  # 48 B boards, 24 B masks (2 swap bitboards)
    with pytest.raises(Exception):
        object.__setattr__  # frozen
        c.rows = 3


def test_codec_against_oracle(E):
    from oracle.oracle import Oracle
    for R, T in ((5, 2), (6, 4), (9, 6), (12, 7), (16, 8)):
        c, o = E.BoardConfig(seed=1, rows=R, columns=R, types=T), Oracle(R, R, T)
        for a in range(c.action_space):
            assert c.decode(a) == o.decode(a)


def test_argument_errors_are_reported_not_raised_in_c(E):
    N = E._native
    L = N.lib()
    cfg = N.Config()
    assert L.ecg_config_init(C.byref(cfg), 9, 8, 6) < 0 and b"square" in L.ecg_last_error()
    assert L.ecg_config_init(C.byref(cfg), 3, 3, 6) < 0 and L.ecg_config_init(C.byref(cfg), 17, 17, 6) < 0
    for size in range(4, 17):  # every square size the reference can run is built in
        assert L.ecg_config_init(C.byref(cfg), size, size, 6) == 0 and cfg.action_space == size * (size - 1) * 2
    assert L.ecg_config_init(C.byref(cfg), 9, 9, 12) < 0
    with pytest.raises(E.EcgError):
        E.BoardConfig(seed=1, rows=9, columns=9, types=0)
    assert L.ecg_config_init(C.byref(cfg), 9, 9, 6) == 0
    rf = N.Refill()
    rf.mode = 7
    io = N.StepIO()
    assert L.ecg_step(C.byref(cfg), C.byref(rf), C.byref(io), 10, None) < 0
    rf.mode = N.REFILL_PHILOX
    assert L.ecg_step(C.byref(cfg), C.byref(rf), C.byref(io), 10, None) < 0 and b"boards_in" in L.ecg_last_error()
    # expanding (board, action) pairs: in/out arrays would race between jobs (read at src_index[i], written at i)
    io.boards_in, io.boards_out, io.actions, io.src_index = 256, 512, 768, 1024  # never dereferenced: checks come first
    io.moves_left = 2048
    assert L.ecg_step(C.byref(cfg), C.byref(rf), C.byref(io), 10, None) < 0 and b"src_index" in L.ecg_last_error()
    io.moves_left = None
    io.boards_out = io.boards_in
    assert L.ecg_step(C.byref(cfg), C.byref(rf), C.byref(io), 10, None) < 0 and b"src_index" in L.ecg_last_error()
    cfg.board_words = 13  # a config not made by ecg_config_init
    assert L.ecg_legal_mask(C.byref(cfg), None, None, 1, None) < 0
    assert L.ecg_boards_bytes(C.byref(E.BoardConfig(seed=1).native), 33) == 64 * 48
    assert L.ecg_masks_bytes(C.byref(E.BoardConfig(seed=1).native), 33) == 64 * 24


def test_no_cpu_fallback(E):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(E.EcgError):
        E.BatchedBoards(E.BoardConfig(seed=1), 4)


def test_product_package_never_uses_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may import, load or link it."""
    pkg = os.path.join(ROOT, "element-crush-gym_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|libecg_oracle|ecgo_|oracle[/\\.]oracle|hostsim", re.M)
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src.replace("tests/hostsim", "")), f


@pytest.mark.parametrize("size", [4, 5, 6, 8, 9, 11, 16])
def test_host_expand_nibbles_matches_the_code_table(E, size):
    """ecg_host_expand_nibbles / the expander pool (host code, no CUDA): 4-bit codes -> uint8 cell values, every code,
    odd and even cell counts, boards below and above the 16-byte SIMD block, ranges cut at every board"""
    import numpy as np
    N = E._native
    L = N.lib()
    types = 11 if size % 2 else 6
    cfg = N.make_config(size, size, types)
    rc, nb = size * size, (size * size + 1) // 2
    lut = np.array(list(range(12)) + [cfg.h_line, cfg.v_line, cfg.bomb, cfg.mega_token], dtype=np.uint8)
    rng = np.random.default_rng(size)
    n = 1003
    nib = rng.integers(0, 256, size=(n, nb), dtype=np.uint8)
    want = np.empty((n, nb * 2), dtype=np.uint8)
    want[:, 0::2] = lut[nib & 15]
    want[:, 1::2] = lut[nib >> 4]
    want = np.ascontiguousarray(want[:, :rc])
    guard = 64
    out = np.full(n * rc + guard, 0xEE, dtype=np.uint8)
    assert L.ecg_host_expand_nibbles(C.byref(cfg), nib.ctypes.data, out.ctypes.data, n) == 0
    assert np.array_equal(out[:n * rc].reshape(n, rc), want) and (out[n * rc:] == 0xEE).all()
    x = L.ecg_host_expander_create(5)
    assert x
    try:
        out[:] = 0xEE
        for lo, hi, split in ((0, 1, 1), (1, 2, 7), (2, 500, 13), (500, n, 64)):
            assert L.ecg_host_expander_submit(x, C.byref(cfg), nib[lo:].ctypes.data, out[lo * rc:].ctypes.data, hi - lo,
                                              None, split) == 0
        assert L.ecg_host_expander_wait(x) == 0
        assert np.array_equal(out[:n * rc].reshape(n, rc), want) and (out[n * rc:] == 0xEE).all()
        assert L.ecg_host_expander_wait(x) == 0  # nothing pending: returns at once
    finally:
        L.ecg_host_expander_destroy(x)
    assert L.ecg_host_expand_nibbles(C.byref(cfg), None, out.ctypes.data, 1) == -1
    assert b"NULL" in L.ecg_last_error()

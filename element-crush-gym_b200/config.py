"""BoardConfig: the host-side mirror of match3tile/boardConfig.py:5-69 (same field names, same derived
constants, same action codec), plus the handle of the engine's native ecg_config."""
from __future__ import annotations

import random
from dataclasses import dataclass, field

from . import _native


@dataclass(frozen=True)
class BoardConfig:
    seed: int = None
    rows: int = 9
    columns: int = 9
    types: int = 6

    shape: tuple = field(init=False)
    action_space: int = field(init=False)
    actions: dict = field(init=False, repr=False, compare=False)
    type_mask: int = field(init=False)
    special_type_mask: int = field(init=False)
    h_line: int = field(init=False)
    v_line: int = field(init=False)
    bomb: int = field(init=False)
    mega_token: int = field(init=False)
    native: object = field(init=False, repr=False, compare=False)

    def __post_init__(self):
        put = lambda k, v: object.__setattr__(self, k, v)  # noqa: E731  (frozen dataclass)
        # boardConfig.py:34 replaces a falsy seed (None or 0) by a random one
        put("seed", self.seed or random.randrange(1, 2 ** 31 - 1))
        nat = _native.make_config(self.rows, self.columns, self.types)  # raises for unsupported shapes
        put("native", nat)
        put("shape", (self.rows, self.columns))
        put("action_space", nat.action_space)
        put("type_mask", nat.type_mask)
        put("special_type_mask", nat.special_type_mask)
        put("h_line", nat.h_line)
        put("v_line", nat.v_line)
        put("bomb", nat.bomb)
        put("mega_token", nat.mega_token)
        put("actions", {a: self.decode(a) for a in range(nat.action_space)})

    def decode(self, action):
        """action -> ((row1, col1), (row2, col2)); the first cell is the upper/left one ("source"),
        the second the lower/right one ("target").  Each board row owns 2*columns-1 actions: its
        columns-1 horizontal swaps, then its columns vertical swaps (boardConfig.py:45-59)."""
        per_row = 2 * self.columns - 1
        row, k = divmod(int(action), per_row)
        if k < self.columns - 1:
            return (row, k), (row, k + 1)
        col = k - (self.columns - 1)
        return (row, col), (row + 1, col)

    def encode(self, tile1, tile2):
        """inverse of decode for two adjacent cells in either order (boardConfig.py:61-69)"""
        (r1, c1), (r2, c2) = tile1, tile2
        assert (c1 == c2 and abs(r1 - r2) == 1) or (r1 == r2 and abs(c1 - c2) == 1), \
            "source and target must be adjacent"
        per_row = 2 * self.columns - 1
        vertical = self.columns - 1 if c1 == c2 else 0
        return min(r1, r2) * per_row + vertical + min(c1, c2)

    @property
    def special_values(self):
        return (self.h_line, self.v_line, self.bomb, self.mega_token)

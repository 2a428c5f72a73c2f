"""Board geometry and the board-record layout shared by host code, kernels and tests.

The reference fixes geometry at compile time (`/root/reference/src/cpp/engine/board.h:22-24`:
``rows_``, ``cols_``, ``invalid_area``).  Here it is a value: ``Geometry(R, IA)``.

Board record (one per game, `include/fpc.h`): ``R*R`` piece bytes in the reference's own
``Piece`` bit layout (`engine/board.h:101-104`: ``present<<7 | color<<5 | type<<2``, empty =
``0x18``), then ``turn`` (1 B), the four ``CastlingRights`` bytes (`engine/board.h:290-291`:
``0x80 | ks<<6 | qs<<5``), the four king squares (``R*R`` = captured, `engine/board.h:193`),
zero padding to a multiple of 16 B.  14x14 -> 208 B, 8x8 -> 80 B.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING, NO_PIECE = range(7)
RED, BLUE, YELLOW, GREEN = range(4)
RED_YELLOW, BLUE_GREEN = 0, 1
IN_PROGRESS, WIN_RY, WIN_BG, STALEMATE = range(4)
EMPTY = 0x18
NUM_STATE_CHANNELS = 24


def piece_byte(color: int, ptype: int) -> int:
    return 0x80 | (color << 5) | (ptype << 2)


def rights_byte(kingside: bool, queenside: bool) -> int:
    return 0x80 | (int(bool(kingside)) << 6) | (int(bool(queenside)) << 5)


@dataclass(frozen=True)
class Geometry:
    R: int
    IA: int

    @property
    def nsq(self) -> int:
        return self.R * self.R

    @property
    def record_bytes(self) -> int:
        return ((self.nsq + 12 + 15) // 16) * 16

    @property
    def off_turn(self) -> int:
        return self.nsq

    @property
    def off_rights(self) -> int:
        return self.nsq + 1

    @property
    def off_king(self) -> int:
        return self.nsq + 5

    @property
    def num_action_channels(self) -> int:
        # src/cpp/board.cpp:11 -- 4*rows + 4*cols + 8 (only 8*(R-1)+8 are reachable, move.cpp:18-20)
        return 8 * self.R + 8

    @property
    def action_space_size(self) -> int:
        return self.num_action_channels * self.nsq

    @property
    def state_space_size(self) -> int:
        return NUM_STATE_CHANNELS * self.nsq

    def is_legal_location(self, row: int, col: int) -> bool:
        """`engine/board.h:647-654`."""
        R, IA = self.R, self.IA
        if row < 0 or row >= R or col < 0 or col >= R:
            return False
        corner_col = col < IA or col > R - 1 - IA
        if row < IA and corner_col:
            return False
        if row > R - 1 - IA and corner_col:
            return False
        return True

    def empty_record(self) -> np.ndarray:
        rec = np.zeros(self.record_bytes, dtype=np.uint8)
        rec[: self.nsq] = EMPTY
        rec[self.off_rights : self.off_rights + 4] = rights_byte(False, False)
        rec[self.off_king : self.off_king + 4] = self.nsq
        return rec


GEOMETRIES = {14: Geometry(14, 3), 8: Geometry(8, 2), 10: Geometry(10, 2), 13: Geometry(13, 3)}

"""FEN -> board record.  Mirrors `/root/reference/src/py/fen_parser.py:104-170`
(`parse_board_args_from_fen`) and the five start positions of
`/root/reference/src/py/start_fens.py`.

Format: fields split on ``-``; field 0 = side to move (R/B/Y/G); fields 2 and 3 = kingside /
queenside castling availability ``a,b,c,d`` in colour order; last field = placement, rows split on
``/`` (top row first), cells on ``,``: ``rP`` = red pawn ..., ``x`` = one skipped cell, an integer =
that many empty cells.

The reference's Python path computes the castling rights and then drops them
(`fen_parser.py:137-140,170`), so a board built the reference's way has no rights.  ``castling``
selects: ``False`` = reference-faithful (all rights off), ``True`` = honour the FEN fields.
"""
from __future__ import annotations

import numpy as np

from .geometry import (BISHOP, BLUE, GREEN, KING, KNIGHT, PAWN, QUEEN, RED, ROOK, YELLOW,
                       GEOMETRIES, Geometry, piece_byte, rights_byte)

_COLORS = {"r": RED, "b": BLUE, "y": YELLOW, "g": GREEN}
_TYPES = {"P": PAWN, "R": ROOK, "N": KNIGHT, "B": BISHOP, "K": KING, "Q": QUEEN}
_TURNS = {"R": RED, "B": BLUE, "Y": YELLOW, "G": GREEN}


def _back_rank(color: str, order: str) -> str:
    return ",".join(color + t for t in order)


def _start_fen(R: int, IA: int, top: str, left: list[str], bottom: str, right: list[str]) -> str:
    """Compose a symmetric four-army start position; armies are given as piece-letter strings."""
    inner = R - 2 * IA
    rows = []
    pad = ",".join(["x"] * IA)
    rows.append(f"{pad},{_back_rank('y', top)},{pad}")
    rows.append(f"{pad},{_back_rank('y', 'P' * inner)},{pad}")
    for _ in range(IA - 2):
        rows.append(f"{pad},{inner},{pad}")
    for lp, rp in zip(left, right):
        rows.append(f"b{lp},bP,{R - 4},gP,g{rp}")
    for _ in range(IA - 2):
        rows.append(f"{pad},{inner},{pad}")
    rows.append(f"{pad},{_back_rank('r', 'P' * inner)},{pad}")
    rows.append(f"{pad},{_back_rank('r', bottom)},{pad}")
    return "R-0,0,0,0-1,1,1,1-1,1,1,1-0,0,0,0-0-" + "/".join(rows)


# start_fens.py:1-16 -- the chess.com four-player teams set-up on the 14x14 cut-corner board.
STANDARD = _start_fen(14, 3, "RNBKQBNR", list("RNBQKBNR"), "RNBQKBNR", list("RNBKQBNR"))
# start_fens.py:18-32
THIRTEEN = _start_fen(13, 3, "RNBKQBN", list("RNBQKBN"), "RNBQKBN", list("RNBKQBN"))
# start_fens.py:34-45
TEN = _start_fen(10, 2, "RNKQBR", list("RNQKBR"), "RNQKBR", list("RNKQBR"))
# start_fens.py:47-56
EIGHT = _start_fen(8, 2, "RKQR", list("RQKR"), "RQKR", list("RKQR"))
# start_fens.py:58-67 -- the reference's default (`four_player_chess_board.py:18`); irregular.
EIGHT_SIMPLE = ("R-0,0,0,0-1,1,1,1-1,1,1,1-0,0,0,0-0-"
                "x,x,yR,yK,x,yR,x,x/x,x,x,yP,yP,yP,x,x/x,x,4,x,x/x,x,4,gP,gK/"
                "bK,bP,4,x,x/x,x,4,x,x/x,x,rP,rP,rP,x,x,x/x,x,rR,x,rK,rR,x,x")

START_FENS = {"STANDARD": (STANDARD, 14), "THIRTEEN": (THIRTEEN, 13), "TEN": (TEN, 10),
              "EIGHT": (EIGHT, 8), "EIGHT_SIMPLE": (EIGHT_SIMPLE, 8)}


def record_from_fen(fen: str, geom: Geometry | int, castling: bool = False) -> np.ndarray:
    if not isinstance(geom, Geometry):
        geom = GEOMETRIES[int(geom)]
    parts = fen.replace("\n", "").replace(" ", "").split("-")
    if len(parts[0]) != 1 or parts[0] not in _TURNS:
        raise ValueError("Invalid player character in FEN string")
    rec = geom.empty_record()
    rec[geom.off_turn] = _TURNS[parts[0]]

    def availability(s: str, what: str) -> list[bool]:
        f = s.split(",")
        if len(f) != 4:
            raise ValueError(f"Invalid {what} castling availability in FEN string")
        return [x == "1" for x in f]

    ks = availability(parts[2], "kingside")
    qs = availability(parts[3], "queenside")
    if castling:
        for c in range(4):
            rec[geom.off_rights + c] = rights_byte(ks[c], qs[c])

    rows = parts[-1].split("/")
    if len(rows) > geom.R:
        raise ValueError("Too many rows in piece placement")
    for row, row_str in enumerate(rows):
        col = 0
        for cell in row_str.split(","):
            if not cell:
                raise ValueError("Empty column string in piece placement")
            ch = cell[0]
            if ch in _COLORS:
                if len(cell) != 2 or cell[1] not in _TYPES:
                    raise ValueError("Piece placement string for player must be of length 2")
                if col >= geom.R:
                    raise ValueError("Piece placement outside the board")
                color, ptype = _COLORS[ch], _TYPES[cell[1]]
                rec[row * geom.R + col] = piece_byte(color, ptype)
                if ptype == KING:
                    rec[geom.off_king + color] = row * geom.R + col
                col += 1
            elif ch == "x":
                col += 1
            else:
                try:
                    n = int(cell)
                except ValueError:
                    n = 0
                if n <= 0:
                    raise ValueError("Invalid number of empty spaces in piece placement")
                col += n
    return rec


def start_record(name: str = "STANDARD", castling: bool = False) -> np.ndarray:
    fen, R = START_FENS[name]
    return record_from_fen(fen, R, castling=castling)

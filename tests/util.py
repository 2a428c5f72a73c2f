"""Shared helpers for the test-suite (test infrastructure)."""
from __future__ import annotations

import functools

import numpy as np

from alphazero_4_player_chess_b200.fen import START_FENS, start_record
from alphazero_4_player_chess_b200.geometry import GEOMETRIES
from oracle.port import Oracle

SEED = 0x5EED


@functools.lru_cache(maxsize=None)
def oracle_for(R: int) -> Oracle:
    return Oracle(R, GEOMETRIES[R].IA)


@functools.lru_cache(maxsize=None)
def playout_positions(name: str, castling: bool, n_games: int, max_plies: int, first_game: int = 0):
    """Positions visited by the oracle's deterministic playouts: dict of stacked arrays."""
    _, R = START_FENS[name]
    o = oracle_for(R)
    start = start_record(name, castling=castling)
    recs, games, plies = [], [], []
    for g in range(first_game, first_game + n_games):
        p = o.playout(start, SEED, g, max_plies)
        recs.append(p["recs"])
        games.append(np.full(p["n"], g))
        plies.append(np.arange(p["n"]))
    return dict(R=R, recs=np.concatenate(recs), game=np.concatenate(games), ply=np.concatenate(plies))


def mixed_positions(name: str, n_positions: int, max_plies: int = 400):
    """At least n_positions positions, half from castling-on and half from castling-off games."""
    out = []
    per = 0
    g = 8
    while per < n_positions:
        a = playout_positions(name, True, g, max_plies)
        b = playout_positions(name, False, g, max_plies, first_game=100000)
        per = len(a["recs"]) + len(b["recs"])
        out = [a, b]
        g *= 2
    recs = np.concatenate([x["recs"] for x in out])[:n_positions]
    return np.ascontiguousarray(recs)


def oracle_legal_lists(R: int, recs: np.ndarray):
    o = oracle_for(R)
    return [o.legal_moves(r) for r in recs]


def castling_positions(R: int = 14):
    """Hand-made castling cases on the 14x14 board (engine/board.cpp:343-465), every colour to move: clean
    castling both sides, a piece in between, the king in check, the crossed square attacked, the destination
    attacked (pseudo-legal, rejected by the legal filter), the partner's rook on the rook square, rights off."""
    g = GEOMETRIES[R]
    base = start_record("STANDARD", castling=True)
    out = []
    for color in range(4):
        rec = base.copy()
        rec[g.off_turn] = color
        # strip every knight, bishop, queen and pawn: kings and rooks only
        for sq in range(g.nsq):
            p = int(rec[sq])
            if p & 0x80 and ((p >> 2) & 7) not in (3, 5):
                rec[sq] = 0x18
        out.append(rec.copy())                                   # clean: both castles available
        ksq = int(rec[g.off_king + color])
        kr, kc = divmod(ksq, R)
        # unit step along the back rank towards the kingside rook
        step = {0: (0, 1), 1: (1, 0), 2: (0, -1), 3: (-1, 0)}[color]
        inward = {0: (-1, 0), 1: (0, 1), 2: (1, 0), 3: (0, -1)}[color]  # away from the own edge

        def sq_at(k, depth=0):
            return (kr + step[0] * k + inward[0] * depth) * R + (kc + step[1] * k + inward[1] * depth)

        enemy_rook = 0x80 | (((color + 1) & 3) << 5) | (3 << 2)
        partner_rook = 0x80 | (((color + 2) & 3) << 5) | (3 << 2)
        for k in (1, 2, -1, -2, -3):                               # a blocker on each between square
            r2 = rec.copy()
            r2[sq_at(k)] = 0x80 | (color << 5) | (1 << 2)
            out.append(r2)
        for k in (0, 1, 2, -1, -2):                                # an enemy rook staring down the file at from / crossed / dest
            r2 = rec.copy()
            r2[sq_at(k, 5)] = enemy_rook
            out.append(r2)
        r2 = rec.copy()                                            # partner's rook on the kingside rook square
        r2[sq_at(3)] = partner_rook
        out.append(r2)
        r2 = rec.copy()                                            # enemy rook there instead
        r2[sq_at(3)] = enemy_rook
        out.append(r2)
        for bits in (0x80, 0x80 | 0x40, 0x80 | 0x20):              # rights: none, kingside only, queenside only
            r2 = rec.copy()
            r2[g.off_rights + color] = bits
            out.append(r2)
    return np.stack(out)


def random_positions(R: int, n: int, seed: int = 7):
    """Random (mostly unreachable) positions: all four kings (one of them missing now and then), up to 10 random
    pieces per colour on random on-board squares, pawns only on squares before their promotion line, random side to
    move and random castling rights.  A differential test bed that does not depend on what playouts reach."""
    g = GEOMETRIES[R]
    rng = np.random.default_rng(seed)
    squares = [sq for sq in range(g.nsq) if g.is_legal_location(sq // R, sq % R)]
    out = []

    def pawn_ok(color, r, c):  # strictly before the promotion line, in the pawn's direction of travel
        return {0: r > R // 4, 2: r < 3 * R // 4, 1: c < 3 * R // 4, 3: c > R // 4}[color]

    for _ in range(n):
        rec = g.empty_record()
        rec[g.off_turn] = rng.integers(0, 4)
        free = list(rng.permutation(squares))
        for color in range(4):
            if rng.random() > 0.1:
                sq = int(free.pop())
                rec[sq] = 0x80 | (color << 5) | (5 << 2)
                rec[g.off_king + color] = sq
            for _ in range(int(rng.integers(0, 11))):
                ptype = int(rng.choice([0, 0, 0, 1, 2, 3, 4]))
                sq = int(free.pop())
                if ptype == 0 and not pawn_ok(color, sq // R, sq % R):
                    ptype = 1
                rec[sq] = 0x80 | (color << 5) | (ptype << 2)
            rec[g.off_rights + color] = 0x80 | (int(rng.integers(0, 2)) << 6) | (int(rng.integers(0, 2)) << 5)
        out.append(rec)
    return np.stack(out)

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

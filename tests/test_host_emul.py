"""CPU: the product's rules kernel body (csrc/fpc_rules.cuh: rules_warp, the code every warp of rules_kernel runs)
compiled for the host and run under a fibre-per-lane warp emulator (tests/host_emul/), against the oracle.  This
checks the whole warp choreography -- line tables, pins, checks, generation, castling, canonical order, the bit sets
of the dense outputs, the playout pick and make-move -- in the GPU-less build container; the -m gpu tests run the
same code on a B200."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from alphazero_4_player_chess_b200.fen import START_FENS, start_record
from alphazero_4_player_chess_b200.geometry import GEOMETRIES
from tests.util import SEED, castling_positions, oracle_for

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "alphazero_4_player_chess_b200", "csrc")


class Params(C.Structure):
    """struct fpc::ObserveParams (csrc/fpc_rules.cuh)."""
    _fields_ = [("boards_in", C.c_void_p), ("boards_out", C.c_void_p), ("n", C.c_int), ("need_movegen", C.c_int),
                ("moves", C.c_void_p), ("flat", C.c_void_p), ("counts", C.c_void_p), ("status", C.c_void_p),
                ("k", C.c_void_p), ("k_all", C.c_int), ("cells", C.c_void_p), ("flats", C.c_void_p),
                ("inc_planes", C.c_void_p), ("inc_mask", C.c_void_p), ("playout", C.c_int), ("seed", C.c_uint64), ("game", C.c_void_p),
                ("ply", C.c_void_p), ("start", C.c_void_p), ("max_plies", C.c_int), ("game_stride", C.c_uint64),
                ("chosen", C.c_void_p), ("counters", C.c_void_p)]


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "host_emul", "emul.cpp")
    deps = [src, os.path.join(HERE, "host_emul", "warp_emul.h"), os.path.join(CSRC, "fpc_device.cuh"),
            os.path.join(CSRC, "fpc_rules.cuh")]
    so = os.path.join(HERE, "host_emul", "libemul.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-DFPC_HOST_EMUL",
                               "-I/usr/local/cuda/include", "-I" + CSRC, "-o", so, src])
    L = C.CDLL(so)
    L.emul_rules.argtypes = [C.c_int, C.POINTER(Params)]
    return L


class Harness:
    """One batch of positions through rules_warp: every output the kernel can produce."""

    def __init__(self, L, R):
        self.L, self.R, self.g = L, R, GEOMETRIES[R]
        cf, cs, ff, fs = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        assert L.emul_strides(C.byref(cf), C.byref(cs), C.byref(ff), C.byref(fs)) == 0
        self.cell_first, self.cell_stride, self.flat_first, self.flat_stride = cf.value, cs.value, ff.value, fs.value

    def observe(self, recs, k_all=-1, playout=None, dense=True):
        """playout: None or dict(seed, games, plies, start, max_plies)."""
        recs = np.ascontiguousarray(recs, dtype=np.uint8)
        n = len(recs)
        out = dict(moves=np.zeros((n, 300), np.uint64), flat=np.zeros((n, 300), np.int32), counts=np.zeros(n, np.int32),
                   status=np.zeros(n, np.int32), cells=np.full((n, self.cell_stride), 0xDEAD, np.uint16),
                   flats=np.full((n, self.flat_stride), 0xDEAD, np.uint16))
        p = Params()
        p.boards_in = recs.ctypes.data
        p.n, p.need_movegen, p.k_all = n, 1, k_all
        for name in ("moves", "flat", "counts", "status"):
            setattr(p, name, out[name].ctypes.data)
        if dense:
            p.cells, p.flats = out["cells"].ctypes.data, out["flats"].ctypes.data
        if playout is not None:
            out["boards"] = recs.copy()
            out["game"] = np.ascontiguousarray(playout["games"], dtype=np.uint64)
            out["ply"] = np.ascontiguousarray(playout["plies"], dtype=np.int32)
            out["chosen"] = np.zeros(n, np.uint64)
            out["counters"] = np.zeros(8, np.uint64)
            self._start = np.ascontiguousarray(playout["start"], dtype=np.uint8)
            p.boards_out = out["boards"].ctypes.data
            p.playout, p.seed, p.max_plies, p.game_stride = 1, playout["seed"], playout["max_plies"], 1000
            p.game, p.ply, p.start = out["game"].ctypes.data, out["ply"].ctypes.data, self._start.ctypes.data
            p.chosen, p.counters = out["chosen"].ctypes.data, out["counters"].ctypes.data
        assert self.L.emul_rules(self.R, C.byref(p)) == 0
        return out

    def dense(self, records, first, n_floats):
        """records of ones -> the 0/1 f32 tensor expand_kernel would write (zero fill, then the ones)."""
        out = np.zeros((len(records), n_floats), np.float32)
        for i, rec in enumerate(records):
            out[i, rec[first: first + int(rec[0])].astype(np.int64)] = 1.0
        return out


def check_positions(h, o, recs, k_all=-1):
    R, g = h.R, h.g
    out = h.observe(recs, k_all=k_all)
    n_check = 0
    for i, rec in enumerate(recs):
        want = o.legal_moves(rec)
        nl = int(out["counts"][i])
        assert nl == len(want), (i, nl, len(want))
        assert np.array_equal(out["moves"][i, :nl], want), i
        assert out["flat"][i, :nl].tolist() == [o.move_flat_index(m) for m in want]
        res, _, kc = o.game_result(rec)
        st = int(out["status"][i])
        turn = int(rec[R * R])
        has_king = any(int(b) == (0x80 | (turn << 5) | (5 << 2)) for b in rec[: R * R])
        if res != 0 and has_king and nl > 0:
            # the reference's order-dependent early-out: IN_PROGRESS + CAN_TAKE_KING here
            assert (st & 3) == 0 and (st & 0x200)
        else:
            assert (st & 3) == res, (i, st, res)
            assert bool(st & 0x200) == kc
        in_check = has_king and o.king_in_check(rec, turn)
        assert bool(st & 0x1000) == in_check, (i, st)
        n_check += in_check
    turns = recs[:, R * R].astype(np.int32)
    k = turns if k_all < 0 else np.full(len(recs), k_all, np.int32)
    planes = h.dense(out["cells"], h.cell_first, g.state_space_size).reshape(-1, 24, R, R)
    assert np.array_equal(planes, o.encode(recs, k))
    mask = h.dense(out["flats"], h.flat_first, g.action_space_size).reshape(-1, g.num_action_channels, R, R)
    assert np.array_equal(mask, o.mask(recs))
    assert (out["flats"][:, 0] == out["counts"]).all()
    return n_check


@pytest.mark.parametrize("name,n_games,max_plies", [("STANDARD", 10, 600), ("THIRTEEN", 5, 300), ("TEN", 6, 300),
                                                    ("EIGHT", 6, 300), ("EIGHT_SIMPLE", 6, 300)])
@pytest.mark.parametrize("castling", [True, False])
def test_rules_warp_on_playouts(emul, name, n_games, max_plies, castling):
    _, R = START_FENS[name]
    o = oracle_for(R)
    h = Harness(emul, R)
    start = start_record(name, castling=castling)
    positions = 0
    for game in range(n_games):
        p = o.playout(start, SEED, game, max_plies)
        recs = p["recs"]
        n = p["n"]
        check_positions(h, o, recs, k_all=-1 if game % 2 == 0 else game % 4)
        # the playout step: pick, make-move, re-seeding, counters
        out = h.observe(recs, playout=dict(seed=SEED, games=np.full(n, game), plies=np.arange(n), start=start,
                                           max_plies=max_plies), dense=False)
        for ply in range(n):
            assert int(out["chosen"][ply]) == int(p["moves"][ply])
            st = int(out["status"][ply])
            if p["result"][ply] == 0 and ply + 1 < max_plies:
                assert not st & 0x800
                assert np.array_equal(out["boards"][ply], recs[ply + 1] if ply + 1 < n else o.make_move(recs[ply], p["moves"][ply])), (game, ply)
                assert out["ply"][ply] == ply + 1 and out["game"][ply] == game
            else:
                assert st & 0x800 and np.array_equal(out["boards"][ply], start)
                assert out["ply"][ply] == 0 and out["game"][ply] == game + 1000
        assert int(out["counters"][0]) == n and int(out["counters"][6]) == int(p["n_legal"].sum())
        positions += n
    assert positions > 100


def random_positions(R, n, seed):
    """Mostly unreachable positions: random pieces on random squares, all four kings present most of the time --
    multiple checks, pins by either opponent, adjacent kings, pawns on every rank."""
    g = GEOMETRIES[R]
    rng = np.random.default_rng(seed)
    valid = [r * R + c for r in range(R) for c in range(R)
             if not ((r < g.IA or r > R - 1 - g.IA) and (c < g.IA or c > R - 1 - g.IA))]
    out = []
    for _ in range(n):
        rec = g.empty_record()
        rec[g.off_turn] = rng.integers(0, 4)
        n_pieces = int(rng.integers(4, 40))
        squares = rng.choice(valid, size=min(n_pieces, len(valid)), replace=False)
        kings = list(range(4)) if rng.random() < 0.9 else list(rng.choice(4, size=3, replace=False))
        for j, sq in enumerate(squares):
            if j < len(kings):
                color, ptype = kings[j], 5
                rec[g.off_king + color] = sq
            else:
                color = int(rng.integers(0, 4))
                ptype = int(rng.choice([0, 0, 0, 1, 2, 3, 4]))
            rec[sq] = 0x80 | (color << 5) | (ptype << 2)
        for c in range(4):
            rec[g.off_rights + c] = 0x80 | (int(rng.integers(0, 4)) << 5)
        out.append(rec)
    return np.stack(out)


@pytest.mark.parametrize("R", [14, 13, 10, 8])
def test_rules_warp_on_random_positions(emul, R):
    o = oracle_for(R)
    h = Harness(emul, R)
    recs = random_positions(R, 700, seed=R)
    n_check = check_positions(h, o, recs)
    assert n_check > 50  # checks, double checks and pins are the point of this test


def test_rules_warp_on_castling_positions(emul):
    o = oracle_for(14)
    h = Harness(emul, 14)
    recs = np.stack(castling_positions(14))
    check_positions(h, o, recs)
    castles = 0
    for rec in recs:
        castles += sum(1 for m in o.legal_moves(rec) if ((int(m) >> 32) & 0xff) != 196)
    assert castles >= 8

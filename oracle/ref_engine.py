"""TEST INFRASTRUCTURE.  ctypes view of oracle/_ref/libref_engine_R<R>.so -- the unmodified
reference rules engine behind oracle/ref_harness.cpp.  Import only from tests/ and bench.py's
cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def available(R: int) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", f"libref_engine_R{R}.so"))


class RefEngine:
    def __init__(self, R: int):
        path = os.path.join(_HERE, "_ref", f"libref_engine_R{R}.so")
        self.lib = lib = C.CDLL(path)
        lib.ref_rows.restype = C.c_int
        lib.ref_invalid_area.restype = C.c_int
        lib.ref_record_bytes.restype = C.c_int
        lib.ref_pseudo_moves.argtypes = [_u8p, _u64p, C.c_int]
        lib.ref_legal_moves.argtypes = [_u8p, _u64p, C.c_int, C.c_void_p]
        lib.ref_game_result.argtypes = [_u8p]
        lib.ref_king_in_check.argtypes = [_u8p, C.c_int]
        lib.ref_is_attacked_by_team.argtypes = [_u8p, C.c_int, C.c_int]
        lib.ref_heuristic.argtypes = [_u8p, C.c_int]
        lib.ref_make_move.argtypes = [_u8p, C.c_uint64, _u8p]
        lib.ref_make_index.argtypes = [_u8p, C.c_int, _u8p]
        lib.ref_move_from_flat.argtypes = [C.c_int]
        lib.ref_move_from_flat.restype = C.c_uint64
        lib.ref_move_flat_index.argtypes = [C.c_uint64]
        lib.ref_perft.argtypes = [_u8p, C.c_int]
        lib.ref_perft.restype = C.c_uint64
        lib.ref_mix.argtypes = [C.c_uint64] * 3
        lib.ref_mix.restype = C.c_uint64
        lib.ref_playout.argtypes = [_u8p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_bench_playout.argtypes = [_u8p, C.c_uint64, C.c_int, C.c_uint64, C.c_int,
                                          C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.ref_bench_playout.restype = C.c_double
        lib.ref_env_create.argtypes = [_u8p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int]
        lib.ref_env_create.restype = C.c_void_p
        lib.ref_env_step.argtypes = [C.c_void_p]
        lib.ref_env_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint64)] * 3
        lib.ref_env_get.argtypes = [C.c_void_p, C.c_int, _u8p, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        lib.ref_env_destroy.argtypes = [C.c_void_p]
        self.R = lib.ref_rows()
        self.IA = lib.ref_invalid_area()
        self.record_bytes = lib.ref_record_bytes()
        assert self.R == R

    def pseudo_moves(self, rec):
        out = np.zeros(300, dtype=np.uint64)
        n = self.lib.ref_pseudo_moves(np.ascontiguousarray(rec), out, 300)
        return out[:n].copy()

    def legal_moves(self, rec, with_after=False):
        out = np.zeros(300, dtype=np.uint64)
        after = np.zeros(self.record_bytes, dtype=np.uint8)
        n = self.lib.ref_legal_moves(np.ascontiguousarray(rec), out, 300, after.ctypes.data)
        return (out[:n].copy(), after) if with_after else out[:n].copy()

    def game_result(self, rec) -> int:
        return self.lib.ref_game_result(np.ascontiguousarray(rec))

    def king_in_check(self, rec, color) -> bool:
        return bool(self.lib.ref_king_in_check(np.ascontiguousarray(rec), color))

    def is_attacked_by_team(self, rec, team, sq) -> bool:
        return bool(self.lib.ref_is_attacked_by_team(np.ascontiguousarray(rec), team, sq))

    def heuristic(self, rec, team) -> int:
        return self.lib.ref_heuristic(np.ascontiguousarray(rec), team)

    def make_move(self, rec, move):
        out = np.zeros(self.record_bytes, dtype=np.uint8)
        rc = self.lib.ref_make_move(np.ascontiguousarray(rec), int(move), out)
        if rc != 0:
            raise RuntimeError("reference MakeMove raised")
        return out

    def make_index(self, rec, flat):
        out = np.zeros(self.record_bytes, dtype=np.uint8)
        rc = self.lib.ref_make_index(np.ascontiguousarray(rec), int(flat), out)
        if rc != 0:
            raise RuntimeError("reference MakeMove(index) raised")
        return out

    def move_from_flat(self, flat) -> int:
        return int(self.lib.ref_move_from_flat(int(flat)))

    def move_flat_index(self, move) -> int:
        return self.lib.ref_move_flat_index(int(move))

    def perft(self, rec, depth) -> int:
        return int(self.lib.ref_perft(np.ascontiguousarray(rec), depth))

    def mix(self, seed, game, ply) -> int:
        return int(self.lib.ref_mix(seed, game, ply))

    def playout(self, start, seed, game, max_plies):
        recs = np.zeros((max_plies, self.record_bytes), dtype=np.uint8)
        n_legal = np.zeros(max_plies, dtype=np.int32)
        result = np.zeros(max_plies, dtype=np.int32)
        result_ref = np.zeros(max_plies, dtype=np.int32)
        moves = np.zeros(max_plies, dtype=np.uint64)
        n = self.lib.ref_playout(np.ascontiguousarray(start), seed, game, max_plies,
                                 recs.ctypes.data, n_legal.ctypes.data, result.ctypes.data,
                                 result_ref.ctypes.data, moves.ctypes.data)
        return dict(n=n, recs=recs[:n], n_legal=n_legal[:n], result=result[:n],
                    result_ref=result_ref[:n], moves=moves[:n])

    def bench_playout(self, start, seed, n_threads, min_positions, max_plies):
        pos = C.c_uint64(0)
        chk = C.c_uint64(0)
        rate = self.lib.ref_bench_playout(np.ascontiguousarray(start), seed, n_threads,
                                          min_positions, max_plies, C.byref(pos), C.byref(chk))
        return dict(positions_per_s=rate, positions=pos.value, checksum=chk.value)


class RefEnv:
    """configs[1] on the host: n_slots games resident as reference Board objects, one ply per step."""

    def __init__(self, engine: RefEngine, start, n_slots, seed, first_game=0, stride=None, max_plies=2048,
                 n_threads=1):
        self.e = engine
        self.n = n_slots
        self.h = engine.lib.ref_env_create(np.ascontiguousarray(start), n_slots, seed, first_game,
                                           n_slots if stride is None else stride, max_plies, n_threads)

    def step(self):
        self.e.lib.ref_env_step(self.h)

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.e.lib.ref_env_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(positions=a.value, finished=b.value, sum_legal=c.value)

    def get(self, slot):
        rec = np.zeros(self.e.record_bytes, dtype=np.uint8)
        g, p = C.c_uint64(), C.c_int()
        self.e.lib.ref_env_get(self.h, slot, rec, C.byref(g), C.byref(p))
        return rec, g.value, p.value

    def close(self):
        if self.h:
            self.e.lib.ref_env_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

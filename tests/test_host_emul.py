"""CPU: the product's device rules header (csrc/fpc_device.cuh) compiled for the host and walked
sequentially (tests/host_emul/emul.cpp), against the oracle.  This checks the rules code the kernels
run -- generation, castling, legal filter, result, make, canonical order -- in the GPU-less build
container; the warp choreography is covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from alphazero_4_player_chess_b200.fen import START_FENS, start_record
from tests.util import SEED, oracle_for

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "host_emul", "emul.cpp")
    hdr = os.path.join(ROOT, "alphazero_4_player_chess_b200", "csrc", "fpc_device.cuh")
    so = os.path.join(HERE, "host_emul", "libemul.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-I/usr/local/cuda/include",
                               "-I" + os.path.dirname(hdr), "-o", so, src])
    L = C.CDLL(so)
    L.emul_step.argtypes = [C.c_int, _u8p, C.c_uint64, C.c_uint64, C.c_uint64, _u8p, _u64p, C.POINTER(C.c_int),
                            C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    return L


@pytest.mark.parametrize("name,n_games,max_plies", [("STANDARD", 12, 500), ("THIRTEEN", 6, 300), ("TEN", 8, 300),
                                                    ("EIGHT", 8, 300), ("EIGHT_SIMPLE", 8, 300)])
@pytest.mark.parametrize("castling", [True, False])
def test_device_rules_on_host(emul, name, n_games, max_plies, castling):
    _, R = START_FENS[name]
    o = oracle_for(R)
    start = start_record(name, castling=castling)
    positions = 0
    for game in range(n_games):
        p = o.playout(start, SEED, game, max_plies)
        for ply in range(p["n"]):
            rec = np.ascontiguousarray(p["recs"][ply])
            out = np.zeros_like(rec)
            moves = np.zeros(300, dtype=np.uint64)
            nl, st, npseudo, chosen = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
            assert emul.emul_step(R, rec, SEED, game, ply, out, moves, C.byref(nl), C.byref(st), C.byref(chosen),
                                  C.byref(npseudo)) == 0
            want = o.legal_moves(rec)
            assert nl.value == len(want) and np.array_equal(moves[: nl.value], want), (game, ply)
            assert npseudo.value == len(o.pseudo_moves(rec))
            res, _, kc = o.game_result(rec)
            assert (st.value & 3) == res and bool(st.value & 0x200) == kc
            assert chosen.value == int(p["moves"][ply])
            if res == 0 and ply + 1 < p["n"]:
                assert np.array_equal(out, p["recs"][ply + 1]), (game, ply)
            positions += 1
    assert positions > 100

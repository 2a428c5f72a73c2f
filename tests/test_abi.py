"""CPU: the C-ABI library loads and exports every symbol include/fpc.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fpc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fpc_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ["fpc_observe", "fpc_make_moves", "fpc_make_index", "fpc_encode", "fpc_playout_step",
                 "fpc_host_observe", "fpc_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from alphazero_4_player_chess_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/fpc.h but not exported"
    assert set(_lib.SIGNATURES) == set(declared_symbols())


def test_static_queries_and_move_index_map_need_no_gpu():
    from alphazero_4_player_chess_b200 import _lib
    from tests.util import oracle_for
    L = _lib.lib()
    assert L.fpc_record_bytes(14) == 208 and L.fpc_record_bytes(8) == 80
    assert L.fpc_num_action_channels(14) == 120 and L.fpc_action_space_size(14) == 23520
    assert L.fpc_state_space_size(8) == 1536 and L.fpc_invalid_area(14) == 3 and L.fpc_invalid_area(8) == 2
    assert L.fpc_supported(12) == 0
    for R in (14, 8, 10, 13):
        o = oracle_for(R)
        for flat in range(0, (8 * (R - 1) + 8) * R * R, 37):
            m = L.fpc_move_from_flat(R, flat)
            assert m == o.move_from_flat(flat)
            assert L.fpc_move_flat_index(R, m) == o.move_flat_index(m)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from alphazero_4_player_chess_b200 import _lib
    L = _lib.lib()
    assert not L.fpc_ctx_create(0, 14, 16)
    assert b"no usable CUDA device" in L.fpc_last_error()


def test_fen_loader_matches_the_python_mirror_and_the_reference():
    """fpc_record_from_fen (C++, host-only) == fen.py == the reference's own parser output (golden)."""
    import numpy as np
    from alphazero_4_player_chess_b200 import _lib
    from alphazero_4_player_chess_b200.fen import START_FENS, record_from_fen
    L = _lib.lib()
    golden = os.path.join(ROOT, "tests", "golden")
    for name, (fen, R) in START_FENS.items():
        for castling in (0, 1):
            out = np.zeros(L.fpc_record_bytes(R), dtype=np.uint8)
            assert L.fpc_record_from_fen(R, fen.encode(), castling, out.ctypes.data) == 0, L.fpc_last_error()
            assert np.array_equal(out, record_from_fen(fen, R, castling=bool(castling))), (name, castling)
        path = os.path.join(golden, f"binding_R{R}.npz")
        if os.path.exists(path):
            z = np.load(path)
            if f"start_{name}" in z.files:
                out = np.zeros(L.fpc_record_bytes(R), dtype=np.uint8)
                assert L.fpc_record_from_fen(R, fen.encode(), 0, out.ctypes.data) == 0
                assert np.array_equal(out, z[f"start_{name}"]), name
    out = np.zeros(208, dtype=np.uint8)
    assert L.fpc_record_from_fen(14, b"Q-0,0,0,0-1,1,1,1-1,1,1,1-0,0,0,0-0-x", 0, out.ctypes.data) == _lib.FPC_ERR_ARG
    assert b"Invalid player character" in L.fpc_last_error()
    assert L.fpc_record_from_fen(14, b"R-0,0,0,0-1,1,1-1,1,1,1-0,0,0,0-0-x", 0, out.ctypes.data) == _lib.FPC_ERR_ARG
    assert L.fpc_record_from_fen(14, b"R-0,0,0,0-1,1,1,1-1,1,1,1-0,0,0,0-0-rZ", 0, out.ctypes.data) == _lib.FPC_ERR_ARG

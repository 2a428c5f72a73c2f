"""-m gpu: the CUDA path (through the C-ABI of include/fpc.h) against the oracle, bit-exact.

Integer/byte/index work: equality.  The f32 planes and masks hold only 0.0 / 1.0, so they are
compared with exact equality too (tolerance 0)."""
import ctypes as C

import numpy as np
import pytest
import torch

from alphazero_4_player_chess_b200 import _lib
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import START_FENS, start_record
from alphazero_4_player_chess_b200.geometry import GEOMETRIES
from tests.util import SEED, mixed_positions, oracle_for

pytestmark = pytest.mark.gpu

PERFT = {  # SURVEY 8c known answers, reproduced by oracle/_ref in tests/test_oracle_vs_ref.py
    "STANDARD": [20, 395, 7800, 152050, 3450730],
    "EIGHT_SIMPLE": [14, 66, 887, 4086, 58416, 258524],
    "EIGHT": [10, 83, 677, 4828, 41693],
    "TEN": [14, 215, 2856, 41416],
}


def gpu_perft(name, depth, castling):
    from alphazero_4_player_chess_b200.perft import perft
    _, R = START_FENS[name]
    return perft(R, start_record(name, castling=castling), depth, chunk=50000)


@pytest.mark.parametrize("name", list(PERFT))
@pytest.mark.parametrize("castling", [False, True])
def test_perft_known_answers(name, castling):
    want = PERFT[name]
    depth = len(want) if name != "STANDARD" else 5
    assert gpu_perft(name, depth, castling) == want[:depth]


@pytest.mark.parametrize("name,n", [("STANDARD", 16384), ("EIGHT_SIMPLE", 4096), ("EIGHT", 2048), ("TEN", 4096),
                                    ("THIRTEEN", 2048)])
def test_observe_matches_oracle(name, n):
    """configs[2]: legal lists, results, planes and masks for synthetic playout positions."""
    _, R = START_FENS[name]
    g = GEOMETRIES[R]
    o = oracle_for(R)
    recs = mixed_positions(name, n)
    n = recs.shape[0]
    env = BatchedEnv(R, n)
    env.load(recs)
    env.observe(planes=True, mask=True, moves=True, flat=True, k=-1)
    torch.cuda.synchronize()
    counts = env.counts.cpu().numpy()
    status = env.status.cpu().numpy()
    moves = env.moves_buffer().cpu().numpy().view(np.uint64)
    flat = env.flat_buffer().cpu().numpy()
    for i in range(n):
        want = o.legal_moves(recs[i])
        assert counts[i] == len(want), i
        assert (moves[i, : len(want)] == want).all(), i
        assert [o.move_flat_index(m) for m in want] == flat[i, : len(want)].tolist(), i
        res, nl, kc = o.game_result(recs[i])
        assert (status[i] & 3) == res and bool(status[i] & _lib.STATUS_CAN_TAKE_KING) == kc, i
        assert not (status[i] & _lib.STATUS_OVERFLOW)
    # dense outputs, chunked (the f32 mask is 94 KB per position at 14x14)
    turns = recs[:, g.off_turn].astype(np.int32)
    planes = env.planes_buffer()
    mask = env.mask_buffer()
    for lo in range(0, n, 1024):
        hi = min(n, lo + 1024)
        assert np.array_equal(planes[lo:hi].cpu().numpy(), o.encode(recs[lo:hi], turns[lo:hi]))
        assert np.array_equal(mask[lo:hi].cpu().numpy(), o.mask(recs[lo:hi]))
    # the reference rotates a whole batch by the colour of states[0]: fixed k, and a k tensor
    for k in range(4):
        got = env.encode(k=k)[:512].cpu().numpy()
        assert np.array_equal(got, o.encode(recs[:512], k))
    kt = torch.arange(n, dtype=torch.int32, device="cuda") % 4
    got = env.encode(k=kt)[:512].cpu().numpy()
    assert np.array_equal(got, o.encode(recs[:512], kt[:512].cpu().numpy()))


@pytest.mark.parametrize("name", ["STANDARD", "EIGHT_SIMPLE", "TEN"])
def test_make_full_and_index_match_oracle(name):
    _, R = START_FENS[name]
    o = oracle_for(R)
    L = _lib.lib()
    recs = mixed_positions(name, 600)[::3]
    parents, full, idx = [], [], []
    for r in recs:
        for m in o.legal_moves(r):
            parents.append(r)
            full.append(m)
            idx.append(o.move_flat_index(m))
    parents = np.stack(parents)
    n = len(full)
    d_par = torch.as_tensor(parents).cuda()
    d_mv = torch.as_tensor(np.array(full, dtype=np.uint64).view(np.int64)).cuda()
    d_ix = torch.as_tensor(np.array(idx, dtype=np.int32)).cuda()
    out = torch.empty_like(d_par)
    err = torch.zeros(n, dtype=torch.int32, device="cuda")
    _lib.check(L.fpc_make_moves(R, d_par.data_ptr(), d_mv.data_ptr(), n, out.data_ptr(), err.data_ptr(), None))
    got = out.cpu().numpy()
    assert int(err.abs().sum().item()) == 0
    for i in range(n):
        assert np.array_equal(got[i], o.make_move(parents[i], full[i])), i
    _lib.check(L.fpc_make_index(R, d_par.data_ptr(), d_ix.data_ptr(), n, out.data_ptr(), err.data_ptr(), None))
    got = out.cpu().numpy()
    assert int(err.abs().sum().item()) == 0
    for i in range(n):
        assert np.array_equal(got[i], o.make_index(parents[i], idx[i])), i


@pytest.mark.parametrize("R,n", [(14, 3000), (13, 1000), (10, 1000), (8, 1500)])
def test_random_positions(R, n):
    """Random, mostly unreachable positions (tests/util.random_positions; the oracle is pinned on the same generator
    against the unmodified engine in tests/test_oracle_vs_ref.py): legal lists, results and flags, planes, masks,
    and the board after the first and last legal move (full and index-built)."""
    from tests.util import random_positions
    g = GEOMETRIES[R]
    L = _lib.lib()
    o = oracle_for(R)
    recs = random_positions(R, n, seed=11)
    env = BatchedEnv(R, n)
    env.load(recs)
    env.observe(planes=True, mask=True, moves=True, flat=True)
    torch.cuda.synchronize()
    counts, status = env.counts.cpu().numpy(), env.status.cpu().numpy()
    moves = env.moves_buffer().cpu().numpy().view(np.uint64)
    flat = env.flat_buffer().cpu().numpy()
    parents, mvs, idx = [], [], []
    n_king_takers = 0
    for i in range(n):
        want = o.legal_moves(recs[i])
        assert counts[i] == len(want) and (moves[i, : len(want)] == want).all(), i
        res, nl, kc = o.game_result(recs[i])
        assert (status[i] & 3) == res and bool(status[i] & _lib.STATUS_CAN_TAKE_KING) == kc, i
        n_king_takers += kc
        for m in ([want[0], want[-1]] if len(want) else []):
            parents.append(recs[i])
            mvs.append(m)
            idx.append(o.move_flat_index(m))
    assert n_king_takers > 0  # positions where a king can be captured do occur in this set
    turns = recs[:, g.off_turn].astype(np.int32)
    assert np.array_equal(env.planes_buffer().cpu().numpy(), o.encode(recs, turns))
    assert np.array_equal(env.mask_buffer().cpu().numpy(), o.mask(recs))
    d_par = torch.as_tensor(np.stack(parents)).cuda()
    d_mv = torch.as_tensor(np.array(mvs, dtype=np.uint64).view(np.int64)).cuda()
    d_ix = torch.as_tensor(np.array(idx, dtype=np.int32)).cuda()
    out, err = torch.empty_like(d_par), torch.zeros(len(mvs), dtype=torch.int32, device="cuda")
    _lib.check(L.fpc_make_moves(R, d_par.data_ptr(), d_mv.data_ptr(), len(mvs), out.data_ptr(), err.data_ptr(), None))
    got = out.cpu().numpy()
    assert not err.any()
    for i in range(len(mvs)):
        assert np.array_equal(got[i], o.make_move(parents[i], mvs[i])), i
    _lib.check(L.fpc_make_index(R, d_par.data_ptr(), d_ix.data_ptr(), len(mvs), out.data_ptr(), err.data_ptr(), None))
    got = out.cpu().numpy()
    assert not err.any()
    for i in range(len(mvs)):
        assert np.array_equal(got[i], o.make_index(parents[i], idx[i])), i


def test_hand_made_castling_cases():
    """Castling details (engine/board.cpp:343-465) on hand-made 14x14 positions: legal lists, flags and the boards
    after every legal move (rook relocation, rights update) against the oracle."""
    from tests.util import castling_positions
    R = 14
    L = _lib.lib()
    o = oracle_for(R)
    recs = castling_positions(R)
    n = len(recs)
    env = BatchedEnv(R, n)
    env.load(recs)
    env.observe(planes=True, mask=True, moves=True)
    torch.cuda.synchronize()
    counts = env.counts.cpu().numpy()
    moves = env.moves_buffer().cpu().numpy().view(np.uint64)
    parents, mvs, castles = [], [], 0
    for i in range(n):
        want = o.legal_moves(recs[i])
        assert counts[i] == len(want) and (moves[i, : len(want)] == want).all(), i
        for m in want:
            parents.append(recs[i])
            mvs.append(m)
            castles += ((int(m) >> 32) & 0xff) != R * R
    assert castles >= 16
    assert np.array_equal(env.mask_buffer().cpu().numpy(), o.mask(recs))
    d_par = torch.as_tensor(np.stack(parents)).cuda()
    d_mv = torch.as_tensor(np.array(mvs, dtype=np.uint64).view(np.int64)).cuda()
    out, err = torch.empty_like(d_par), torch.zeros(len(mvs), dtype=torch.int32, device="cuda")
    _lib.check(L.fpc_make_moves(R, d_par.data_ptr(), d_mv.data_ptr(), len(mvs), out.data_ptr(), err.data_ptr(), None))
    got = out.cpu().numpy()
    assert not err.any()
    for i in range(len(mvs)):
        assert np.array_equal(got[i], o.make_move(parents[i], mvs[i])), i


def test_make_missing_piece_reports_error():
    R = 14
    g = GEOMETRIES[R]
    L = _lib.lib()
    rec = start_record("STANDARD")
    empty_sq = 6 * R + 6
    mv = np.array([empty_sq | ((empty_sq + 1) << 8) | (0x18 << 16) | (6 << 24) | (g.nsq << 32) | (g.nsq << 40)],
                  dtype=np.uint64)
    d_in = torch.as_tensor(rec).cuda().unsqueeze(0).contiguous()
    d_mv = torch.as_tensor(mv.view(np.int64)).cuda()
    out = torch.empty_like(d_in)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(L.fpc_make_moves(R, d_in.data_ptr(), d_mv.data_ptr(), 1, out.data_ptr(), err.data_ptr(), None))
    assert err.item() == _lib.FPC_ERR_MOVE  # "piece missing for move" engine/board.cpp:1046-1054


@pytest.mark.parametrize("name,n_games,steps", [("STANDARD", 512, 260), ("EIGHT_SIMPLE", 256, 200), ("TEN", 128, 200)])
@pytest.mark.parametrize("castling", [True, False])
def test_playout_replays_oracle_games(name, n_games, steps, castling):
    """configs[1]: the fused playout kernel plays the very games the oracle plays."""
    _, R = START_FENS[name]
    o = oracle_for(R)
    start = start_record(name, castling=castling)
    max_plies = 120
    env = BatchedEnv(R, n_games)
    env.reset_playout(start)
    boards, counts, status, chosen, games, plies = [], [], [], [], [], []
    for _ in range(steps):
        games.append(env.game.cpu().numpy().copy())
        plies.append(env.ply.cpu().numpy().copy())
        boards.append(env.boards.cpu().numpy().copy())
        env.playout_step(seed=SEED, max_plies=max_plies, planes=False, mask=False, chosen=True)
        counts.append(env.counts.cpu().numpy().copy())
        status.append(env.status.cpu().numpy().copy())
        chosen.append(env.chosen.cpu().numpy().view(np.uint64).copy())
    total = 0
    for slot in range(0, n_games, 7):
        cache = {}
        for t in range(steps):
            gid, ply = int(games[t][slot]), int(plies[t][slot])
            if gid not in cache:
                cache[gid] = o.playout(start, SEED, gid, max_plies)
            p = cache[gid]
            assert ply < p["n"], (slot, t, gid, ply)
            assert np.array_equal(boards[t][slot], p["recs"][ply]), (slot, t)
            assert counts[t][slot] == p["n_legal"][ply]
            assert (status[t][slot] & 3) == p["result"][ply]
            assert chosen[t][slot] == p["moves"][ply]
            finished = p["result"][ply] != 0 or ply + 1 >= max_plies
            assert bool(status[t][slot] & _lib.STATUS_FINISHED) == finished
            total += 1
    assert int(env.counters[0].item()) == n_games * steps
    assert int(env.counters[6].item()) == int(np.sum(counts))
    assert total > 0


def test_host_api_matches_device_api():
    R = 14
    L = _lib.lib()
    o = oracle_for(R)
    recs = mixed_positions("STANDARD", 1000)
    n = recs.shape[0]
    ctx = L.fpc_ctx_create(0, R, n)
    assert ctx, L.fpc_last_error()
    try:
        moves = np.zeros((n, 300), dtype=np.uint64)
        flat = np.zeros((n, 300), dtype=np.int32)
        counts = np.zeros(n, dtype=np.int32)
        status = np.zeros(n, dtype=np.int32)
        planes = np.zeros((n, 24, R, R), dtype=np.float32)
        mask = np.zeros((n, 8 * R + 8, R, R), dtype=np.float32)
        _lib.check(L.fpc_host_observe(ctx, recs.ctypes.data, n, moves.ctypes.data, flat.ctypes.data,
                                      counts.ctypes.data, status.ctypes.data, planes.ctypes.data, None, -1,
                                      mask.ctypes.data, None))
        for i in range(n):
            want = o.legal_moves(recs[i])
            assert counts[i] == len(want) and (moves[i, : len(want)] == want).all()
        turns = recs[:, GEOMETRIES[R].off_turn].astype(np.int32)
        assert np.array_equal(planes, o.encode(recs, turns))
        assert np.array_equal(mask, o.mask(recs))
        # make through host buffers: first legal move of every position that has one
        has = counts > 0
        par = np.ascontiguousarray(recs[has])
        mv = np.ascontiguousarray(moves[has, 0])
        out = np.zeros_like(par)
        err = np.zeros(len(par), dtype=np.int32)
        _lib.check(L.fpc_host_make_moves(ctx, par.ctypes.data, mv.ctypes.data, len(par), out.ctypes.data,
                                         err.ctypes.data))
        assert not err.any()
        for i in range(len(par)):
            assert np.array_equal(out[i], o.make_move(par[i], mv[i]))
        ix = np.ascontiguousarray(flat[has, 0])
        _lib.check(L.fpc_host_make_index(ctx, par.ctypes.data, ix.ctypes.data, len(par), out.ctypes.data,
                                         err.ctypes.data))
        for i in range(len(par)):
            assert np.array_equal(out[i], o.make_index(par[i], ix[i]))
    finally:
        L.fpc_ctx_destroy(ctx)


@pytest.mark.parametrize("pinned", [True, False])
def test_host_playout_step_zero_copy_and_staged(pinned):
    """fpc_host_playout_step: pinned host buffers are read/written by the kernel directly (zero-copy),
    pageable ones go through staging copies; both must replay the device-resident env exactly."""
    R, n, steps = 14, 512, 60
    L = _lib.lib()
    g = GEOMETRIES[R]
    start = start_record("STANDARD", castling=True)
    env = BatchedEnv(R, n)
    env.reset_playout(start)
    mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    h_boards = mk(torch.from_numpy(np.broadcast_to(start, (n, g.record_bytes)).copy()))
    h_game = mk(torch.arange(n, dtype=torch.int64))
    h_ply = mk(torch.zeros(n, dtype=torch.int32))
    h_counts = mk(torch.zeros(n, dtype=torch.int32))
    h_status = mk(torch.zeros(n, dtype=torch.int32))
    h_start = torch.from_numpy(start.copy())
    planes = torch.empty((n, 24, R, R), dtype=torch.float32, device="cuda")
    mask = torch.empty((n, g.num_action_channels, R, R), dtype=torch.float32, device="cuda")
    ctx = L.fpc_ctx_create(0, R, n)
    assert ctx, L.fpc_last_error()
    try:
        for step in range(steps):
            flags = _lib.FLAG_ASYNC_DENSE if step % 2 else 0
            _lib.check(L.fpc_host_playout_step(ctx, h_boards.data_ptr(), n, SEED, h_game.data_ptr(), h_ply.data_ptr(),
                                               h_start.data_ptr(), 40, n, h_counts.data_ptr(), h_status.data_ptr(),
                                               planes.data_ptr(), -1, mask.data_ptr(), flags))
            env.playout_step(seed=SEED, max_plies=40, planes=True, mask=True)
            torch.cuda.synchronize()
            assert np.array_equal(h_boards.numpy(), env.boards.cpu().numpy()), step
            assert np.array_equal(h_game.numpy(), env.game.cpu().numpy())
            assert np.array_equal(h_ply.numpy(), env.ply.cpu().numpy())
            assert np.array_equal(h_counts.numpy(), env.counts.cpu().numpy())
            assert np.array_equal(h_status.numpy(), env.status.cpu().numpy())
        _lib.check(L.fpc_ctx_sync(ctx))
        assert torch.equal(planes, env.planes_buffer()) and torch.equal(mask, env.mask_buffer())
    finally:
        L.fpc_ctx_destroy(ctx)


def test_edge_cases_empty_ragged_terminal():
    R = 14
    g = GEOMETRIES[R]
    L = _lib.lib()
    o = oracle_for(R)
    # n = 0 is a no-op
    _lib.check(L.fpc_observe(R, None, 0, None, None, None, None, None, None, -1, None, 0, None))
    # unsupported geometry
    assert L.fpc_observe(12, None, 0, None, None, None, None, None, None, -1, None, 0, None) == _lib.FPC_ERR_ARG
    recs = []
    # (a) mover has no king -> other team wins, no moves (engine/board.cpp:852-856, 895-899)
    r = start_record("STANDARD")
    r[13 * R + 7] = 0x18
    r[g.off_king + 0] = g.nsq
    recs.append(r)
    # (b) bare kings: legal moves only for the king
    r = g.empty_record()
    for color, sq in enumerate([13 * R + 7, 7 * R + 0, 0 * R + 6, 6 * R + 13]):
        r[sq] = 0x80 | (color << 5) | (5 << 2)
        r[g.off_king + color] = sq
    recs.append(r)
    # (c) stalemate: red king boxed in by its own pawns' blockers, not attacked
    r = g.empty_record()
    r[13 * R + 3] = 0x80 | (0 << 5) | (5 << 2)      # red king in the corner of its back rank
    r[12 * R + 3] = 0x80 | (1 << 5) | (3 << 2)      # blue rook guards... placed so king cannot move
    r[11 * R + 4] = 0x80 | (3 << 5) | (4 << 2)      # green queen covers the rest
    r[0 * R + 6] = 0x80 | (2 << 5) | (5 << 2)
    r[7 * R + 0] = 0x80 | (1 << 5) | (5 << 2)
    r[6 * R + 13] = 0x80 | (3 << 5) | (5 << 2)
    recs.append(r)
    # (d) empty board, every colour to move
    for t in range(4):
        r = g.empty_record()
        r[g.off_turn] = t
        recs.append(r)
    recs = np.stack(recs)
    n = recs.shape[0]  # 7: not a multiple of the 4 warps per block
    env = BatchedEnv(R, n)
    env.load(recs)
    env.observe(planes=True, mask=True, moves=True)
    torch.cuda.synchronize()
    counts, status = env.counts.cpu().numpy(), env.status.cpu().numpy()
    moves = env.moves_buffer().cpu().numpy().view(np.uint64)
    for i in range(n):
        want = o.legal_moves(recs[i])
        res, nl, kc = o.game_result(recs[i])
        assert counts[i] == len(want) and (moves[i, : len(want)] == want).all(), i
        assert (status[i] & 3) == res, i
    assert np.array_equal(env.mask_buffer().cpu().numpy(), o.mask(recs))
    assert np.array_equal(env.planes_buffer().cpu().numpy(), o.encode(recs, recs[:, g.off_turn].astype(np.int32)))


def test_async_dense_pipeline_matches_sync():
    """FPC_FLAG_ASYNC_DENSE: the expansion of step t overlaps the rules of step t+1; after fpc_join the
    dense tensors are those of the last step, bit-identical to the synchronous path and the oracle."""
    R, n = 14, 1024
    g = GEOMETRIES[R]
    o = oracle_for(R)
    start = start_record("STANDARD", castling=True)
    a, b = BatchedEnv(R, n), BatchedEnv(R, n)
    a.reset_playout(start)
    b.reset_playout(start)
    for step in range(40):
        before = a.boards.clone()
        a.playout_step(seed=SEED, planes=True, mask=True, async_dense=True)
        b.playout_step(seed=SEED, planes=True, mask=True)
    a.join()
    torch.cuda.synchronize()
    assert torch.equal(a.boards, b.boards)
    assert torch.equal(a.planes_buffer(), b.planes_buffer()) and torch.equal(a.mask_buffer(), b.mask_buffer())
    recs = before[:256].cpu().numpy()
    assert np.array_equal(a.planes_buffer()[:256].cpu().numpy(), o.encode(recs, recs[:, g.off_turn].astype(np.int32)))
    assert np.array_equal(a.mask_buffer()[:256].cpu().numpy(), o.mask(recs))


@pytest.mark.parametrize("name,R,n", [("STANDARD", 14, 777), ("THIRTEEN", 13, 130), ("EIGHT_SIMPLE", 8, 257)])
def test_incremental_dense_update_equals_full_rewrite(name, R, n):
    """FPC_FLAG_INCREMENTAL: resident planes / mask tensors updated in place (previous ones cleared, new ones set)
    stay bit-identical to a full rewrite, step after step, across reloads, rotations and plane-only calls; an
    output set the library has no record of falls back to the full rewrite."""
    start = start_record(name, castling=True)
    a, b = BatchedEnv(R, n), BatchedEnv(R, n)
    a.reset_playout(start)
    b.reset_playout(start)
    a.planes_buffer().fill_(7.0)  # garbage: the first incremental call has no record and must rewrite everything
    a.mask_buffer().fill_(7.0)
    for step in range(45):
        a.playout_step(seed=SEED, max_plies=30, planes=True, mask=True, incremental=True)
        b.playout_step(seed=SEED, max_plies=30, planes=True, mask=True)
        if step % 5 == 0 or step > 40:
            torch.cuda.synchronize()
            assert torch.equal(a.planes_buffer(), b.planes_buffer()), step
            assert torch.equal(a.mask_buffer(), b.mask_buffer()), step
    assert torch.equal(a.boards, b.boards)
    # observe with a fixed rotation, then new boards in the same buffers
    for k in (1, -1, 3):
        a.observe(planes=True, mask=True, k=k, incremental=True)
        b.observe(planes=True, mask=True, k=k)
        assert torch.equal(a.planes_buffer(), b.planes_buffer()) and torch.equal(a.mask_buffer(), b.mask_buffer())
    fresh = mixed_positions(name, n)
    a.load(fresh)
    b.load(fresh)
    a.observe(planes=True, mask=True, incremental=True)
    b.observe(planes=True, mask=True)
    assert torch.equal(a.planes_buffer(), b.planes_buffer()) and torch.equal(a.mask_buffer(), b.mask_buffer())
    # interleaving a full rewrite keeps the record valid
    a.playout_step(seed=SEED, planes=True, mask=True)
    a.playout_step(seed=SEED, planes=True, mask=True, incremental=True)
    b.playout_step(seed=SEED, planes=True, mask=True)
    b.playout_step(seed=SEED, planes=True, mask=True)
    torch.cuda.synchronize()
    assert torch.equal(a.planes_buffer(), b.planes_buffer()) and torch.equal(a.mask_buffer(), b.mask_buffer())
    # ... and so do calls that rewrite only ONE of the two tensors between incremental calls (encode(), a planes-only
    # observe with another rotation, a mask-only observe): the record of each tensor is kept on its own
    for variant in range(4):
        if variant == 0:
            a.encode(k=2)
        elif variant == 1:
            a.observe(planes=True, mask=False, k=1)
        elif variant == 2:
            a.observe(planes=False, mask=True)
        else:
            a.observe(planes=True, mask=False, k=3, incremental=True)
        a.playout_step(seed=SEED, planes=True, mask=True, incremental=True)
        b.playout_step(seed=SEED, planes=True, mask=True)
        torch.cuda.synchronize()
        assert torch.equal(a.planes_buffer(), b.planes_buffer()), variant
        assert torch.equal(a.mask_buffer(), b.mask_buffer()), variant
    # a write the library did not make is declared with invalidate_dense(): the next call rewrites everything
    a.planes_buffer().fill_(3.0)
    a.mask_buffer().fill_(3.0)
    a.invalidate_dense()
    a.observe(planes=True, mask=True, incremental=True)
    b.observe(planes=True, mask=True)
    assert torch.equal(a.planes_buffer(), b.planes_buffer()) and torch.equal(a.mask_buffer(), b.mask_buffer())


def test_incremental_never_trusts_a_reused_address():
    """Tensors freed and re-allocated at the same address: the new owner's first FPC_FLAG_INCREMENTAL call must be a
    full rewrite.  Content tracking lives in explicit fpc_dense_track handles, never in pointer identity; the plain
    (handle-less) entry points ignore the flag altogether."""
    R, n = 14, 300
    start = start_record("STANDARD", castling=True)
    ref = BatchedEnv(R, n)
    ref.reset_playout(start)
    ref.observe(planes=True, mask=True)
    torch.cuda.synchronize()
    want_p, want_m = ref.planes_buffer().clone(), ref.mask_buffer().clone()
    ptrs = set()
    for round_ in range(3):
        env = BatchedEnv(R, n)
        env.reset_playout(start)
        for _ in range(3):
            env.playout_step(seed=SEED, planes=True, mask=True, incremental=True)
        torch.cuda.synchronize()
        ptrs.add((env.planes_buffer().data_ptr(), env.mask_buffer().data_ptr()))
        env.close()
        del env  # torch's caching allocator hands the same blocks to the next BatchedEnv
        env = BatchedEnv(R, n)
        env.reset_playout(start)
        env.planes_buffer().fill_(9.0)  # whatever the previous owner left, plus garbage
        env.mask_buffer().fill_(9.0)
        ptrs.add((env.planes_buffer().data_ptr(), env.mask_buffer().data_ptr()))
        env.observe(planes=True, mask=True, incremental=True)
        torch.cuda.synchronize()
        assert torch.equal(env.planes_buffer(), want_p) and torch.equal(env.mask_buffer(), want_m), round_
        # the handle-less C entry point with the flag set: a full rewrite as well
        env.planes_buffer().fill_(5.0)
        env.mask_buffer().fill_(5.0)
        _lib.check(env.L.fpc_observe(R, env.boards.data_ptr(), n, None, None, env.counts.data_ptr(), env.status.data_ptr(),
                                     env.planes_buffer().data_ptr(), None, -1, env.mask_buffer().data_ptr(),
                                     _lib.FLAG_INCREMENTAL, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert torch.equal(env.planes_buffer(), want_p) and torch.equal(env.mask_buffer(), want_m), round_
        env.close()
        del env
    assert len(ptrs) < 6  # the addresses were indeed reused at least once
    # per-thread state can be dropped and comes back on demand
    _lib.check(ref.L.fpc_shutdown())
    ref.observe(planes=True, mask=True)
    torch.cuda.synchronize()
    assert torch.equal(ref.planes_buffer(), want_p) and torch.equal(ref.mask_buffer(), want_m)


def test_million_position_playout_checksum():
    """BASELINE.json configs[1] at full size: 4,096 games played to the end (or 300 plies) on the device, > 1 M
    positions; the order-independent checksum over every position's (n_legal, result, move played) must equal
    the oracle's over the same games (oracle/fpc_oracle.c fpo_playout_checksum, 8 host threads)."""
    from concurrent.futures import ThreadPoolExecutor
    R, n, max_plies = 14, 4096, 300
    o = oracle_for(R)
    start = start_record("STANDARD", castling=True)
    env = BatchedEnv(R, n)
    env.reset_playout(start)
    total = torch.zeros((), dtype=torch.int64, device="cuda")
    positions = torch.zeros((), dtype=torch.int64, device="cuda")
    K = torch.tensor(0x9E3779B97F4A7C15 - (1 << 64), dtype=torch.int64, device="cuda")
    for _ in range(max_plies):
        first = env.game < n  # slots still playing their first game (re-seeded slots get ids >= n)
        env.playout_step(seed=SEED, max_plies=max_plies, planes=False, mask=False, chosen=True)
        v = env.counts.long() * 4 + (env.status.long() & 3) + (((env.chosen * K) >> 40) & 0xFFFFFF)
        total += torch.where(first, v, torch.zeros_like(v)).sum()
        positions += first.sum()
    torch.cuda.synchronize()
    assert not bool((env.game < n).any())
    chunks = [(g0, 256) for g0 in range(0, n, 256)]
    with ThreadPoolExecutor(8) as ex:  # ctypes releases the GIL
        parts = list(ex.map(lambda c: o.playout_checksum(start, SEED, c[0], c[1], max_plies), chunks))
    want = sum(p[0] for p in parts) & ((1 << 64) - 1)
    want_pos = sum(p[1] for p in parts)
    assert int(positions.item()) == want_pos and want_pos > 1_000_000
    assert int(total.item()) & ((1 << 64) - 1) == want


def test_dlpack_handoff_is_zero_copy():
    R, n = 8, 64
    env = BatchedEnv(R, n)
    env.load(start_record("EIGHT_SIMPLE"))
    env.observe(planes=True, mask=True)
    for name, buf in (("planes", env.planes_buffer()), ("mask", env.mask_buffer()), ("boards", env.boards)):
        t = torch.from_dlpack(env.dlpack(name))
        assert t.data_ptr() == buf.data_ptr() and t.shape == buf.shape and t.device == buf.device
    torch.cuda.synchronize()
    assert torch.equal(torch.from_dlpack(env.dlpack("planes")), env.planes_buffer())


def test_full_size_properties():
    """BASELINE.json configs[1] size: 4096 games; size-independent invariants of the fused step."""
    R, n = 14, 4096
    g = GEOMETRIES[R]
    env = BatchedEnv(R, n)
    env.reset_playout(start_record("STANDARD", castling=True))
    for step in range(64):
        before = env.boards.clone()
        env.playout_step(seed=SEED, max_plies=2048, planes=True, mask=True)
        if step % 16 == 15:
            planes, mask = env.planes_buffer(), env.mask_buffer()
            pieces = (before[:, : g.nsq] >= 0x80).sum(dim=1)
            assert torch.equal(planes.sum(dim=(1, 2, 3)).long(), pieces.long())
            assert bool(((planes == 0) | (planes == 1)).all()) and bool(((mask == 0) | (mask == 1)).all())
            # distinct (plane, from) pairs <= legal moves (promotions share an index)
            nz = mask.sum(dim=(1, 2, 3)).long()
            assert bool((nz <= env.counts.long()).all()) and bool((nz * 4 >= env.counts.long()).all())
            # the mask only fires on squares holding a piece of the side to move
            turn = before[:, g.off_turn].long()
            sq = before[:, : g.nsq].long()
            own = ((sq >= 0x80) & (((sq >> 5) & 3) == turn[:, None])).view(n, 1, R, R)
            assert bool((mask.amax(dim=1, keepdim=True) <= own.float()).all())
    assert int(env.counters[0].item()) == 64 * n


def test_playout_stepper_is_the_same_call():
    """BatchedEnv.playout_stepper (arguments resolved once) plays the same games as playout_step."""
    start = start_record("STANDARD", castling=True)
    a, b = BatchedEnv(14, 256), BatchedEnv(14, 256)
    a.reset_playout(start)
    b.reset_playout(start)
    step = b.playout_stepper(seed=SEED, max_plies=64, planes=True, mask=True, async_dense=True)
    for _ in range(80):
        a.playout_step(seed=SEED, max_plies=64, planes=True, mask=True, async_dense=True)
        step()
    a.join()
    b.join()
    torch.cuda.synchronize()
    for name in ("boards", "game", "ply", "counts", "status", "counters"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert torch.equal(a.planes_buffer(), b.planes_buffer()) and torch.equal(a.mask_buffer(), b.mask_buffer())

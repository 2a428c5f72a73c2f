"""-m gpu: the viewer queries (fpc_attack_maps + the drop-in's GetSimpleState / GetAttackedSquares* / IsAttackedByPlayer),
the library-owned store with its own DLPack export (fpc_env_*, NativeEnv), and the drop-in's observation cache."""
import os
import sys

import numpy as np
import pytest
import torch

from alphazero_4_player_chess_b200 import _lib, build
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import START_FENS, start_record
from alphazero_4_player_chess_b200.native_env import NativeEnv
from tests.util import SEED, mixed_positions, oracle_for

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def az():
    build.build_binding()
    sys.path.insert(0, build.DROPIN)
    import alphazero_cpp
    return alphazero_cpp


def board_from_record(az, rec, R):
    pieces = {}
    for sq in range(R * R):
        b = int(rec[sq])
        if b & 0x80:
            pieces[az.BoardLocation(sq // R, sq % R)] = az.Piece(az.PlayerColor((b >> 5) & 3), az.PieceType((b >> 2) & 7))
    return az.Board(az.Player(az.PlayerColor(int(rec[R * R]))), pieces)


@pytest.mark.parametrize("name", ["STANDARD", "THIRTEEN", "TEN", "EIGHT"])
def test_attack_maps_match_the_oracle(name):
    _, R = START_FENS[name]
    o = oracle_for(R)
    recs = mixed_positions(name, 1024)
    d = torch.from_numpy(recs).cuda()
    out = torch.zeros((len(recs), R * R), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().fpc_attack_maps(R, d.data_ptr(), len(recs), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    got = out.cpu().numpy()
    for i in range(0, len(recs), 3):
        assert np.array_equal(got[i], o.attack_map(recs[i])), i


@pytest.mark.parametrize("R", [14, 8])
def test_attack_maps_and_dropin_viewer_queries_match_the_reference_fixtures(az, R):
    az.set_board_size(R)
    z = np.load(os.path.join(GOLDEN, f"viewer_R{R}.npz"))
    recs = np.ascontiguousarray(z["recs"])
    d = torch.from_numpy(recs).cuda()
    out = torch.zeros((len(recs), R * R), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().fpc_attack_maps(R, d.data_ptr(), len(recs), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(out.cpu().numpy(), z["attack_map"])
    off = z["simple_pieces_off"]
    for i in range(0, len(recs), 4):
        b = board_from_record(az, recs[i], R)
        want = z["attack_map"][i]
        players = b.GetAttackedSquaresPlayers()
        for c in range(4):
            sqs = sorted(l.GetRow() * R + l.GetCol() for l in players.get(az.PlayerColor(c), []))
            assert sqs == np.nonzero((want >> c) & 1)[0].tolist()
            assert (az.PlayerColor(c) in players) == bool(((want >> c) & 1).any())  # the reference's map has no empty lists
        teams = b.GetAttackedSquaresTeams()
        for t in range(2):
            sqs = sorted(l.GetRow() * R + l.GetCol() for l in teams.get(az.Team(t), []))
            assert sqs == np.nonzero((want >> (4 + t)) & 1)[0].tolist()
        for sq in range(0, R * R, 5):
            for c in range(4):
                assert b.IsAttackedByPlayer(az.BoardLocation(sq // R, sq % R), az.PlayerColor(c)) == bool((want[sq] >> c) & 1)
        st = b.GetSimpleState()
        assert isinstance(st, az.SimpleBoardState) and int(st.turn.GetColor()) == int(z["simple_turn"][i])
        got = sorted((int(pp.GetPiece().GetColor()), int(pp.GetPiece().GetPieceType()),
                      pp.GetLocation().GetRow() * R + pp.GetLocation().GetCol()) for pl in st.pieces for pp in pl)
        assert got == [tuple(x) for x in z["simple_pieces"][off[i]: off[i + 1]].tolist()]
        assert len(st.castlingRights) == 4
        assert {int(k): sorted(l.GetRow() * R + l.GetCol() for l in v) for k, v in st.attackedSquares.items()} == \
               {int(k): sorted(l.GetRow() * R + l.GetCol() for l in v) for k, v in players.items()}


def test_dropin_children_carry_their_observation(az):
    """TakeAction / ExpandNodes children answer GetGameResult and GetLegalMoves from the trip that made them: the answers
    equal those of a board rebuilt from the same record (a cache miss -> its own trip) and the oracle's; SetTurn forgets."""
    R = 14
    az.set_board_size(R)
    o = oracle_for(R)
    recs = mixed_positions("STANDARD", 96)[::4]
    for rec in recs:
        b = board_from_record(az, rec, R)
        rec = np.frombuffer(b.record(), dtype=np.uint8).copy()  # the Python constructor path carries no castling rights
        legal = b.GetLegalMoves()
        assert [m.image() for m in legal] == [int(m) for m in sorted(o.legal_moves(rec), key=lambda m: (o.move_flat_index(int(m)), (int(m) >> 24) & 0xff))]
        for m in legal[:: max(1, len(legal) // 4)]:
            child = b.TakeAction(m)
            crec = np.frombuffer(child.record(), dtype=np.uint8)
            assert np.array_equal(crec, o.make_move(rec, m.image()))
            fresh = board_from_record(az, crec, R)
            assert [x.image() for x in child.GetLegalMoves()] == [x.image() for x in fresh.GetLegalMoves()]
            assert int(child.GetGameResult()) == int(fresh.GetGameResult()) == o.game_result(crec)[0]
        if legal:
            child = b.TakeAction(legal[0])
            before = [x.image() for x in child.GetLegalMoves()]
            turn = int(child.GetTurn().GetColor())
            child.SetTurn(az.Player(az.PlayerColor((turn + 1) % 4)))
            crec = np.frombuffer(child.record(), dtype=np.uint8)
            assert [x.image() for x in child.GetLegalMoves()] == [x.image() for x in board_from_record(az, crec, R).GetLegalMoves()]
            child.SetTurn(az.Player(az.PlayerColor(turn)))
            assert [x.image() for x in child.GetLegalMoves()] == before


def _from_capsule(cap):
    return torch.utils.dlpack.from_dlpack(cap)


def test_native_env_store_and_its_own_dlpack_export():
    """fpc_env_*: the library owns the store and builds the DLManagedTensor itself; torch wraps it without a copy."""
    R, n = 14, 512
    recs = mixed_positions("STANDARD", n)
    o = oracle_for(R)
    env = NativeEnv(R, n, device=0)
    env.load(recs)
    assert np.array_equal(env.boards(), recs)
    env.observe(outputs=("planes", "mask", "moves", "flat"))
    env.sync()
    planes, mask = _from_capsule(env.dlpack("planes")), _from_capsule(env.dlpack("mask"))
    counts, status = _from_capsule(env.dlpack("counts")), _from_capsule(env.dlpack("status"))
    moves, boards = _from_capsule(env.dlpack("moves")), _from_capsule(env.dlpack("boards"))
    assert planes.shape == (n, 24, R, R) and planes.dtype == torch.float32 and planes.is_cuda
    assert mask.shape == (n, 120, R, R) and moves.shape == (n, 300) and moves.dtype == torch.int64
    assert boards.shape == (n, 208) and boards.dtype == torch.uint8 and counts.dtype == torch.int32
    sel = list(range(0, n, 16))
    turns = recs[sel][:, R * R].astype(np.int32)
    assert np.array_equal(planes[sel].cpu().numpy(), o.encode(recs[sel], turns))
    assert np.array_equal(mask[sel].cpu().numpy(), o.mask(recs[sel]))
    for i in sel:
        res, n_legal, _ = o.game_result(recs[i])
        assert int(counts[i]) == n_legal and (int(status[i]) & 3) == res
    # the same tensors against the torch-owned twin
    twin = BatchedEnv(R, n)
    twin.load(recs)
    twin.observe(planes=True, mask=True, moves=True)
    torch.cuda.synchronize()
    assert torch.equal(planes, twin.planes_buffer()) and torch.equal(mask, twin.mask_buffer())
    assert torch.equal(counts, twin.counts)
    # zero-copy: a playout step through the library shows up in the exported tensors
    start = start_record("STANDARD", castling=True)
    env.load(np.tile(start, (n, 1)))
    ply = _from_capsule(env.dlpack("ply"))
    for _ in range(3):
        env.playout_step(start, SEED, outputs=("planes",))
    env.sync()
    assert int(ply.min()) == 3 and not torch.equal(boards.cpu(), torch.from_numpy(np.tile(start, (n, 1))))
    cur = [start.copy() for _ in range(4)]
    for g in range(4):
        for p in range(3):
            cur[g] = o.playout_step(cur[g], SEED, g, p)[1]
        assert np.array_equal(boards[g].cpu().numpy(), cur[g])
    # lifetime: the store outlives fpc_env_destroy while an exported tensor is alive
    keep = planes.clone()
    env.playout_step(start, SEED, outputs=())  # planes untouched by a step that does not ask for them
    env.sync()
    env.close()
    torch.cuda.synchronize()
    assert torch.equal(planes, keep)
    del planes, mask, counts, status, moves, boards, ply
    assert not _lib.lib().fpc_env_dlpack(None, _lib.ENV_PLANES)


def test_dropin_expand_nodes_in_chunks(az):
    """Node.ExpandNodes with more children than one trip of the host-buffer context holds (16,384): 1,000 roots x 20
    moves = 20,000 children made and observed in two chunks; sampled children equal the oracle's make-move and carry the
    right legal moves and result."""
    R = 14
    az.set_board_size(R)
    o = oracle_for(R)
    start = start_record("STANDARD")
    n_roots = 1000
    roots = [board_from_record(az, start, R) for _ in range(n_roots)]
    rec0 = np.frombuffer(roots[0].record(), dtype=np.uint8).copy()
    legal = sorted((int(m) for m in o.legal_moves(rec0)), key=lambda m: (o.move_flat_index(m), (m >> 24) & 0xff))
    flats = [o.move_flat_index(m) for m in legal]
    nodes = [az.Node(3.0, b, visit_count=1) for b in roots]
    nsq = R * R
    nz = [[i, f // nsq, (f % nsq) // R, f % R] for i in range(n_roots) for f in flats]
    values = [1.0 / len(flats)] * len(nz)
    az.Node.ExpandNodes(nodes, torch.zeros(1), nz, values, az.BoardPool(1))
    assert len(nz) > 16384
    for i in (0, 1, 499, 818, 819, 820, 999):  # 819 * 20 = 16,380: the chunk boundary falls inside root 819
        ch = nodes[i].GetChildren()
        assert [c.GetMoveMade().GetFlatIndex() for c in ch] == flats
        for c, f in zip(ch, flats):
            crec = np.frombuffer(c.GetState().record(), dtype=np.uint8)
            assert np.array_equal(crec, o.make_index(rec0, f))
            want = sorted((int(m) for m in o.legal_moves(crec)), key=lambda m: (o.move_flat_index(m), (m >> 24) & 0xff))
            assert [m.image() for m in c.GetState().GetLegalMoves()] == want
            assert int(c.GetState().GetGameResult()) == o.game_result(crec)[0]

"""-m gpu: the CUDA path (C-ABI) against the committed golden fixtures directly -- outputs of the
reference itself (tests/golden/*.npz, generator tests/golden/make_golden.py), no oracle in between."""
import os

import numpy as np
import pytest
import torch

from alphazero_4_player_chess_b200 import _lib
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import START_FENS, start_record
from alphazero_4_player_chess_b200.geometry import GEOMETRIES
from alphazero_4_player_chess_b200.perft import perft

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def unpack(bits, shape):
    return np.unpackbits(bits)[: int(np.prod(shape))].reshape(shape).astype(np.float32)


@pytest.mark.parametrize("name", ["STANDARD", "THIRTEEN", "TEN", "EIGHT", "EIGHT_SIMPLE"])
@pytest.mark.parametrize("tag", ["c0", "c1"])
def test_engine_fixtures_on_the_device(name, tag):
    """Reference engine: perft table, legal lists (canonical order), results, post-move boards, replayed games."""
    _, R = START_FENS[name]
    L = _lib.lib()
    z = np.load(os.path.join(GOLDEN, f"engine_{name}.npz"))
    start = start_record(name, castling=tag == "c1")
    assert perft(R, start, len(z[f"perft_{tag}"])) == z[f"perft_{tag}"].tolist()
    recs = np.ascontiguousarray(z[f"recs_{tag}"])
    n = len(recs)
    env = BatchedEnv(R, n)
    env.load(recs)
    env.observe(planes=False, mask=False, moves=True, flat=True)
    torch.cuda.synchronize()
    counts, status = env.counts.cpu().numpy(), env.status.cpu().numpy()
    moves = env.moves_buffer().cpu().numpy().view(np.uint64)
    legal, off = z[f"legal_{tag}"], z[f"legal_off_{tag}"]
    assert np.array_equal(counts, np.diff(off))
    for i in range(n):
        assert np.array_equal(moves[i, : counts[i]], legal[off[i]: off[i + 1]]), i
    res = status & 3
    assert np.array_equal(res, z[f"result_{tag}"])            # canonical result recorded by the harness
    differs = res != z[f"result_ref_{tag}"]                    # GetGameResult verbatim: only the early-out may differ
    assert np.all((res[differs] == 0) & ((status[differs] & _lib.STATUS_CAN_TAKE_KING) != 0))
    # make(full) and make(index) of the last legal move of every position that has one
    has = counts > 0
    par = torch.as_tensor(recs[has]).cuda()
    last = torch.as_tensor(np.array([moves[i, counts[i] - 1] for i in np.nonzero(has)[0]], dtype=np.uint64).view(np.int64)).cuda()
    flat = torch.as_tensor(np.array([env.flat_buffer()[i, counts[i] - 1].item() for i in np.nonzero(has)[0]], dtype=np.int32)).cuda()
    out, err = torch.empty_like(par), torch.zeros(par.shape[0], dtype=torch.int32, device="cuda")
    _lib.check(L.fpc_make_moves(R, par.data_ptr(), last.data_ptr(), par.shape[0], out.data_ptr(), err.data_ptr(), None))
    assert not err.any() and np.array_equal(out.cpu().numpy(), z[f"after_last_legal_{tag}"][has])
    _lib.check(L.fpc_make_index(R, par.data_ptr(), flat.data_ptr(), par.shape[0], out.data_ptr(), err.data_ptr(), None))
    assert not err.any() and np.array_equal(out.cpu().numpy(), z[f"after_last_legal_index_{tag}"][has])
    # the playout kernel replays the reference's games: boards, n_legal and moves ply by ply
    game = z[f"game_{tag}"]
    games = np.unique(game)
    penv = BatchedEnv(R, len(games))
    penv.reset_playout(start)
    idx = {int(g): np.nonzero(game == g)[0] for g in games}
    for ply in range(max(len(v) for v in idx.values())):
        before = penv.boards.cpu().numpy()
        penv.playout_step(seed=int(z["seed"]), max_plies=int(z["max_plies"]), planes=False, mask=False, chosen=True)
        chosen = penv.chosen.cpu().numpy().view(np.uint64)
        for g in games:
            rows = idx[int(g)]
            if ply < len(rows):
                assert np.array_equal(before[g], recs[rows[ply]]), (g, ply)
                assert chosen[g] == z[f"moves_{tag}"][rows[ply]]


@pytest.mark.parametrize("R", [14, 8])
def test_binding_fixtures_on_the_device(R):
    """Reference binding: encoder planes (per state and states[0]-rotated batches), legal masks, legal indices,
    heuristics, TakeAction(Move(flat_index)) boards."""
    g = GEOMETRIES[R]
    L = _lib.lib()
    z = np.load(os.path.join(GOLDEN, f"binding_R{R}.npz"))
    recs = np.ascontiguousarray(z["recs"])
    n = len(recs)
    env = BatchedEnv(R, n)
    env.load(recs)
    env.observe(planes=True, mask=True, flat=True, k=-1)
    torch.cuda.synchronize()
    assert np.array_equal(env.planes_buffer().cpu().numpy(), unpack(z["planes_own"], (n, 24, R, R)))
    assert np.array_equal(env.mask_buffer().cpu().numpy(), unpack(z["mask"], (n, g.num_action_channels, R, R)))
    counts, flat = env.counts.cpu().numpy(), env.flat_buffer().cpu().numpy()
    off = z["legal_off"]
    assert np.array_equal(counts, np.diff(off))
    for i in range(n):
        assert flat[i, : counts[i]].tolist() == z["legal_flat"][off[i]: off[i + 1]].tolist()
    for k in range(4):
        order = z[f"planes_batch_k{k}_order"]
        sub = BatchedEnv(R, len(order))
        sub.load(recs[order])
        assert np.array_equal(sub.encode(k=k).cpu().numpy(), unpack(z[f"planes_batch_k{k}"], (len(order), 24, R, R)))
    for team in (0, 1):
        r2 = recs.copy()
        r2[:, g.off_turn] = team
        env.load(r2)
        assert env.heuristic().cpu().numpy().tolist() == z["heuristic"][:, team].tolist()
    has = counts > 0
    par = torch.as_tensor(recs[has]).cuda()
    last = torch.as_tensor(np.array([flat[i, counts[i] - 1] for i in np.nonzero(has)[0]], dtype=np.int32)).cuda()
    out, err = torch.empty_like(par), torch.zeros(par.shape[0], dtype=torch.int32, device="cuda")
    _lib.check(L.fpc_make_index(R, par.data_ptr(), last.data_ptr(), par.shape[0], out.data_ptr(), err.data_ptr(), None))
    assert not err.any() and np.array_equal(out.cpu().numpy(), z["after_index"][has])

"""-m gpu: the pybind11 drop-in module `alphazero_cpp` (csrc/binding.cpp -> dropin/) driven through the
reference's own call sequences, against the golden fixtures the reference produced.

`drive_search` issues exactly the calls `src/py/mcts.py:17-89` and
`src/py/four_player_chess_board.py:36-55` issue (Node(C, game, visit_count=1), ChooseLeaf, GetState,
GetEncodedStates, ParseActionspace, GetLegalMoves + GetLegalMovesIndices + index_put_,
BackpropagateNodes, nonzero/tolist, ExpandNodes with a BoardPool); the reference's Python itself cannot
run on the GPU box (no /root/reference there)."""
import os
import sys

import numpy as np
import pytest
import torch

from alphazero_4_player_chess_b200 import build
from tests.golden.fake_net import FakeNet
from tests.util import oracle_for

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


def legal_moves_mask(az, states, device):
    legal = [s.GetLegalMoves() for s in states]
    b, p, r, c = az.Board.GetLegalMovesIndices(legal, sum(len(m) for m in legal))
    mask = torch.zeros((len(states), *az.Board.action_space_dims), dtype=torch.float32, device=device)
    idx = tuple(torch.tensor(v, dtype=torch.int64, device=device) for v in (b, p, r, c))
    mask.index_put_(idx, torch.tensor(1, dtype=torch.float32, device=device))
    return mask


def drive_search(az, games, net, C, sims):
    pool = az.BoardPool(1)
    roots = []
    for g in games:
        root = az.Node(C, g, visit_count=1)
        g.SetRootNode(root)
        roots.append(root)
    live = roots[:]
    for _ in range(sims):
        leaves = []
        for node in live[:]:
            leaf = node.ChooseLeaf()
            if leaf is None:
                live.remove(node)
            else:
                leaves.append(leaf)
        if not leaves:
            continue
        states = [leaf.GetState() for leaf in leaves]
        enc = az.Board.GetEncodedStates(states, "cuda")
        logits, value = net(enc)
        policy = az.Board.ParseActionspace(torch.softmax(logits, dim=1), states[0].GetTurn())
        policy = policy * legal_moves_mask(az, states, "cuda")
        policy = policy / policy.sum(dim=(1, 2, 3), keepdim=True)
        az.Node.BackpropagateNodes(leaves, value.squeeze(1))
        cpu = policy.cpu()
        nz = torch.nonzero(cpu, as_tuple=True)
        az.Node.ExpandNodes(leaves, cpu, torch.nonzero(cpu).tolist(), cpu[nz].tolist(), pool)
    return roots


@pytest.mark.parametrize("R,case", [(8, "a"), (8, "b"), (14, "a")])
def test_reference_call_sequence_reproduces_reference_search(az, R, case):
    az.set_board_size(R)
    z = np.load(os.path.join(GOLDEN, f"mcts_R{R}.npz"))
    games = [board_from_record(az, r, R) for r in z[f"{case}_roots"]]
    roots = drive_search(az, games, FakeNet(R, device="cuda"), 3, int(z[f"{case}_sims"]))
    off = z[f"{case}_child_off"]
    for g, root in enumerate(roots):
        ch = root.GetChildren()
        assert [c.GetMoveMade().GetFlatIndex() for c in ch] == z[f"{case}_child_flat"][off[g]: off[g + 1]].tolist(), g
        assert [c.GetVisitCount() for c in ch] == z[f"{case}_child_visits"][off[g]: off[g + 1]].tolist(), g
        assert root.GetVisitCount() == int(z[f"{case}_root_visits"][g])
        assert games[g].GetRootNode() is not None


@pytest.mark.parametrize("R", [14, 8])
def test_board_methods_against_binding_fixtures(az, R):
    az.set_board_size(R)
    o = oracle_for(R)
    z = np.load(os.path.join(GOLDEN, f"binding_R{R}.npz"))
    recs = z["recs"][:48]
    boards = [board_from_record(az, r, R) for r in recs]
    n, A = len(recs), az.Board.num_action_channels
    for b, r in zip(boards, recs):
        assert np.frombuffer(b.record(), dtype=np.uint8).tolist() == r.tolist()
    planes_own = np.unpackbits(z["planes_own"])[: len(z["recs"]) * 24 * R * R].reshape(-1, 24, R, R)[:n]
    for i, b in enumerate(boards[:16]):
        got = az.Board.GetEncodedStates([b], "cpu")
        assert got.device.type == "cpu" and np.array_equal(got.numpy()[0], planes_own[i].astype(np.float32))
    k = int(recs[0][R * R])
    batch = az.Board.GetEncodedStates(boards, "cuda")
    assert batch.is_cuda and np.array_equal(batch.cpu().numpy(), o.encode(recs, k))
    mask_ref = np.unpackbits(z["mask"])[: len(z["recs"]) * A * R * R].reshape(-1, A, R, R)[:n].astype(np.float32)
    assert np.array_equal(legal_moves_mask(az, boards, "cuda").cpu().numpy(), mask_ref)
    assert np.array_equal(az.Board.LegalMovesMask(boards, "cuda").cpu().numpy(), mask_ref)
    off = z["legal_off"]
    for i, b in enumerate(boards):
        lm = b.GetLegalMoves()
        assert [m.GetFlatIndex() for m in lm] == z["legal_flat"][off[i]: off[i + 1]].tolist()
        res = int(b.GetGameResult())
        if res != int(z["result"][i]):
            assert res == 0
        assert [b.CalculateHeuristic(az.RED_YELLOW), b.CalculateHeuristic(az.BLUE_GREEN)] == z["heuristic"][i].tolist()
        if lm:
            nxt = b.TakeAction(az.Move(lm[-1].GetFlatIndex()))
            assert np.frombuffer(nxt.record(), dtype=np.uint8).tolist() == z["after_index"][i].tolist()
            assert int(nxt.GetTurn().GetColor()) == (int(b.GetTurn().GetColor()) + 1) % 4
    with pytest.raises(RuntimeError):
        az.Board.GetEncodedStates(boards[:1], "tpu")
    with pytest.raises(RuntimeError):  # "piece missing for move"
        boards[0].TakeAction(az.Move(az.BoardLocation(R // 2, R // 2), az.BoardLocation(R // 2, R // 2 + 1)))

"""CPU: the host-side surface of the pybind11 drop-in module `alphazero_cpp` (csrc/binding.cpp) -- names,
value types, constructors, index map, printing -- against the fixtures the reference binding produced.
Rules calls need a CUDA device and must fail loudly without one (no CPU fallback); they are covered by
tests/test_gpu_dropin.py."""
import os
import sys

import numpy as np
import pytest

from alphazero_4_player_chess_b200 import build
from alphazero_4_player_chess_b200.fen import START_FENS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def az():
    build.build_binding()
    sys.path.insert(0, build.DROPIN)
    import alphazero_cpp
    return alphazero_cpp


def board_from_fen(az, fen, R):
    """FEN -> Board through the module's own constructors, the way src/py/fen_parser.py:104-170 does."""
    parts = fen.replace("\n", "").split("-")
    turn = az.Player({"R": az.RED, "B": az.BLUE, "Y": az.YELLOW, "G": az.GREEN}[parts[0]])
    pieces = {}
    for row, row_str in enumerate(parts[-1].split("/")):
        col = 0
        for cell in row_str.split(","):
            if cell[0] in "rbyg":
                color = {"r": az.RED, "b": az.BLUE, "y": az.YELLOW, "g": az.GREEN}[cell[0]]
                ptype = {"P": az.PAWN, "R": az.ROOK, "N": az.KNIGHT, "B": az.BISHOP, "K": az.KING, "Q": az.QUEEN}[cell[1]]
                pieces[az.BoardLocation(row, col)] = az.Piece(az.Player(az.PlayerColor(color)), ptype)
                col += 1
            elif cell == "x":
                col += 1
            else:
                col += int(cell)
    return az.Board(turn, pieces)


def test_module_surface(az):
    for name in ["PieceType", "PlayerColor", "Team", "GameResult", "Player", "Piece", "BoardLocation", "CastlingRights",
                 "PlacedPiece", "Move", "Board", "MemoryEntry", "BoardPool", "Node", "piece_value", "color_value",
                 "RED", "BLUE", "YELLOW", "GREEN", "PAWN", "KNIGHT", "BISHOP", "ROOK", "QUEEN", "KING", "NO_PIECE",
                 "RED_YELLOW", "BLUE_GREEN", "IN_PROGRESS", "WIN_RY", "WIN_BG", "STALEMATE"]:
        assert hasattr(az, name), name
    for name in ["num_state_channels", "state_space_size", "num_action_channels", "action_space_size", "action_space_dims",
                 "state_space_dims", "GetLegalMoves", "TakeAction", "GetGameResult", "GetEncodedStates", "GetEncodedState",
                 "ParseActionspace", "GetLegalMovesIndices", "GetLegalMovesMask", "GetPieces", "GetTurn", "SetTurn",
                 "CalculateHeuristic", "GetRootNode", "SetRootNode", "GetRootState", "SetRootState", "AppendToMemory",
                 "GetMemory", "IsLegalLocation", "nRows", "nCols", "invalidArea", "GetOpponent", "GetOpponentValue",
                 "ChangePerspective", "GetPieceAt", "GetBoardLocation", "IsMoveLegal", "GetSimpleState",
                 "GetAttackedSquaresPlayers", "GetAttackedSquaresTeams", "IsAttackedByPlayer"]:
        assert hasattr(az.Board, name), name
    for name in ["ChooseLeaf", "SelectChild", "Backpropagate", "BackpropagateNodes", "ExpandNodes", "GetChildren",
                 "GetVisitCount", "SetVisitCount", "GetMoveMade", "GetState", "IsExpanded"]:
        assert hasattr(az.Node, name), name
    assert az.piece_value(az.QUEEN) == 4 and az.color_value(az.GREEN) == 3


@pytest.mark.parametrize("R,name", [(14, "STANDARD"), (8, "EIGHT_SIMPLE"), (8, "EIGHT")])
def test_fen_boards_and_piece_strings(az, R, name):
    az.set_board_size(R)
    z = np.load(os.path.join(GOLDEN, f"binding_R{R}.npz"))
    st = z["statics"]
    assert [az.Board.num_state_channels, az.Board.state_space_size, az.Board.num_action_channels, az.Board.action_space_size,
            az.Move.num_queen_moves_per_direction, az.Move.num_queen_moves, az.Move.num_knight_moves,
            az.Board.invalidArea()] == st.tolist()
    assert az.Board.action_space_dims == (int(st[2]), R, R) and az.Board.state_space_dims == (24, R, R)
    b = board_from_fen(az, START_FENS[name][0], R)
    assert np.frombuffer(b.record(), dtype=np.uint8).tolist() == z[f"start_{name}"].tolist()
    # "Red Pawn at e2 (6, 4)": the strings the reference's viewer consumes (state_serializer.py:4-7)
    got = sorted(str(pp) for plist in b.GetPieces() for pp in plist)
    assert got == sorted(z[f"start_{name}_str"].tolist())
    kinds = [[int(pp.GetPiece().GetPieceType()) for pp in plist] for plist in b.GetPieces()]
    for plist in kinds:  # constructor order K, P, N, B, R, Q (engine/board.cpp:1225-1247)
        order = [5, 0, 1, 2, 3, 4]
        assert plist == sorted(plist, key=order.index)
    assert b.GetTurn() == az.Player(az.RED) and int(b.GetTurn().GetTeam()) == 0
    assert az.Board.IsLegalLocation(0, 0) is False and az.Board.IsLegalLocation(R // 2, R // 2) is True


@pytest.mark.parametrize("R", [14, 8])
def test_move_index_map_and_errors(az, R):
    az.set_board_size(R)
    z = np.load(os.path.join(GOLDEN, f"binding_R{R}.npz"))
    for flat, to in zip(z["legal_flat"][:400].tolist(), z["legal_to"][:400].tolist()):
        m = az.Move(flat)
        assert m.GetFlatIndex() == flat
        assert m.To().GetRow() * R + m.To().GetCol() == to
        plane, row, col = m.GetIndex()
        assert az.Move(plane, az.BoardLocation(row, col)).To() == m.To()
    with pytest.raises(RuntimeError):  # unmapped delta (move.cpp:97)
        az.Move(az.BoardLocation(1, 3), az.BoardLocation(6, 4)).GetIndex()
    with pytest.raises(RuntimeError):
        az.Board.GetEncodedStates([], "tpu")
    b = az.Board(az.Player(az.BLUE), {az.BoardLocation(R // 2, 0): az.Piece(az.BLUE, az.KING)})
    with pytest.raises(RuntimeError):
        b.GetPieceAt(R, 0)  # engine/board.h:535
    assert az.Board.GetOpponent(az.GREEN) == az.RED and az.Board.GetOpponentValue(0.5) == -0.5
    root = az.Node(3.0, b, visit_count=1)
    b.SetRootNode(root)
    assert b.GetRootNode().GetVisitCount() == 1 and not root.IsExpanded()
    root.Backpropagate(0.5)
    assert root.GetVisitCount() == 2
    with pytest.raises(RuntimeError):  # "Failed to select a child." (node.cpp:72-75)
        root.SelectChild()
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            b.GetLegalMoves()

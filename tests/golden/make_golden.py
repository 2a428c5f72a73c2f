#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Generates the committed golden fixtures in tests/golden/ by running the
REFERENCE ITSELF in the build container (it cannot travel to the GPU box):

  engine_<NAME>.npz   from the unmodified rules engine (oracle/_ref/libref_engine_R<R>.so, built by
                      oracle/Makefile): perft table, deterministic playouts (positions, legal lists
                      in canonical order, results, moves played, post-move boards).
  binding_R<R>.npz    from the reference's own pybind module `alphazero_cpp` (built by
                      oracle/build_ref_binding.sh) driven through the reference's own Python
                      (`src/py/four_player_chess_board.py`, `fen_parser.py`, `start_fens.py`):
                      start positions as the reference parses them, encoder planes (per-state and
                      states[0]-rotated batches), legal masks, legal flat indices, game results,
                      TakeAction boards for full and index-built moves, the ParseActionspace
                      permutation, heuristics.
  mcts_R<R>.npz       the reference's `MCTS.search` (`src/py/mcts.py`) on the reference binding with
                      the deterministic stand-in network of tests/golden/fake_net.py: per game the root's
                      children (flat action indices, visit counts), the root's visit count and the
                      number of nodes in the tree (the binding exposes neither priors nor value sums).

  viewer_R<R>.npz     the viewer's queries of the reference binding (attacked squares per colour / team,
                      GetSimpleState) on the same positions.

usage:  python tests/golden/make_golden.py            # everything (needs /root/reference)
        python tests/golden/make_golden.py --viewer-only  # only viewer_R*.npz
        python tests/golden/make_golden.py --binding 14   # (internal) one binding geometry
"""
from __future__ import annotations

import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_PY = "/root/reference/src/py"
SEED = 0x5EED

ENGINE_CASES = {  # name -> (n_games, max_plies, perft depth)
    "STANDARD": (3, 240, 4), "THIRTEEN": (2, 160, 3), "TEN": (3, 200, 4), "EIGHT": (4, 200, 5),
    "EIGHT_SIMPLE": (4, 200, 6),
}
BINDING_NAMES = {14: ["STANDARD"], 8: ["EIGHT_SIMPLE", "EIGHT"]}


def make_engine():
    from alphazero_4_player_chess_b200.fen import START_FENS, start_record
    from oracle.ref_engine import RefEngine
    for name, (n_games, max_plies, depth) in ENGINE_CASES.items():
        _, R = START_FENS[name]
        ref = RefEngine(R)
        out = {}
        for castling in (False, True):
            tag = "c1" if castling else "c0"
            start = start_record(name, castling=castling)
            out[f"perft_{tag}"] = np.array([ref.perft(start, d) for d in range(1, depth + 1)], dtype=np.int64)
            recs, n_legal, result, result_ref, moves, game, legal, legal_off = [], [], [], [], [], [], [], [0]
            after_first, after_index = [], []
            for g in range(n_games):
                p = ref.playout(start, SEED, g, max_plies)
                recs.append(p["recs"])
                n_legal.append(p["n_legal"])
                result.append(p["result"])
                result_ref.append(p["result_ref"])
                moves.append(p["moves"])
                game.append(np.full(p["n"], g, dtype=np.int32))
                for rec in p["recs"]:
                    lm = [int(m) for m in ref.legal_moves(rec)]
                    lm.sort(key=lambda m: (ref.move_flat_index(m), (m >> 24) & 0xff))
                    legal.extend(lm)
                    legal_off.append(len(legal))
                    if lm:
                        after_first.append(ref.make_move(rec, lm[-1]))
                        after_index.append(ref.make_index(rec, ref.move_flat_index(lm[-1])))
                    else:
                        after_first.append(rec)
                        after_index.append(rec)
            out[f"recs_{tag}"] = np.concatenate(recs)
            out[f"n_legal_{tag}"] = np.concatenate(n_legal)
            out[f"result_{tag}"] = np.concatenate(result)
            out[f"result_ref_{tag}"] = np.concatenate(result_ref)
            out[f"moves_{tag}"] = np.concatenate(moves)
            out[f"game_{tag}"] = np.concatenate(game)
            out[f"legal_{tag}"] = np.array(legal, dtype=np.uint64)
            out[f"legal_off_{tag}"] = np.array(legal_off, dtype=np.int64)
            out[f"after_last_legal_{tag}"] = np.stack(after_first)
            out[f"after_last_legal_index_{tag}"] = np.stack(after_index)
        out["max_plies"] = np.int64(max_plies)
        out["seed"] = np.int64(SEED)
        path = os.path.join(HERE, f"engine_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path))


# ---- binding (runs in a subprocess with the reference module of one geometry on sys.path) -----------

def _record_of(board, R):
    """Board record (include/fpc.h) of a reference Board, through its bound API only."""
    nsq = R * R
    rec = np.zeros(((nsq + 12 + 15) // 16) * 16, dtype=np.uint8)
    rec[:nsq] = 0x18
    rec[nsq + 1: nsq + 5] = 0x80
    rec[nsq + 5: nsq + 9] = nsq
    rec[nsq] = int(board.GetTurn().GetColor())
    for plist in board.GetPieces():
        for pp in plist:
            loc, piece = pp.GetLocation(), pp.GetPiece()
            sq = loc.GetRow() * R + loc.GetCol()
            color, ptype = int(piece.GetColor()), int(piece.GetPieceType())
            rec[sq] = 0x80 | (color << 5) | (ptype << 2)
            if ptype == 5:
                rec[nsq + 5 + color] = sq
    return rec


def make_binding(R: int):
    import torch
    import alphazero_cpp as az
    import start_fens
    from fen_parser import parse_board_args_from_fen
    from four_player_chess_board import FourPlayerChess
    from mcts import MCTS

    from tests.golden.fake_net import FakeNet

    assert az.Board.nRows() == R
    A = az.Board.num_action_channels
    nsq = R * R
    out = {}

    def board_from_record(rec):
        pieces = {}
        for sq in range(nsq):
            b = int(rec[sq])
            if b & 0x80:
                pieces[az.BoardLocation(sq // R, sq % R)] = az.Piece(az.PlayerColor((b >> 5) & 3),
                                                                    az.PieceType((b >> 2) & 7))
        return FourPlayerChess(az.Player(az.PlayerColor(int(rec[nsq]))), pieces)

    # (1) the start positions exactly as the reference's Python builds them
    for name in BINDING_NAMES[R]:
        fen = getattr(start_fens, name).replace("\n", "")  # as four_player_chess_board.py:18 does
        b = FourPlayerChess(*parse_board_args_from_fen(fen, R))
        out[f"start_{name}"] = _record_of(b, R)
        out[f"start_{name}_str"] = np.array([str(pp) for pl in b.GetPieces() for pp in pl])

    # (2) positions: every 3rd ply of the castling-off engine playouts (the Python path has no rights)
    eng = np.load(os.path.join(HERE, f"engine_{BINDING_NAMES[R][0]}.npz"))
    recs = np.ascontiguousarray(eng["recs_c0"][::3][:256])
    n = len(recs)
    boards = [board_from_record(r) for r in recs]
    for b, r in zip(boards, recs):
        assert np.array_equal(_record_of(b, R), r)
    out["recs"] = recs
    own = torch.cat([az.Board.GetEncodedStates([b], "cpu") for b in boards])
    out["planes_own"] = np.packbits(own.numpy().astype(bool), axis=None)
    assert bool(((own == 0) | (own == 1)).all())
    # whole batches rotated by states[0]'s colour: four batches, one per colour of states[0]
    for k in range(4):
        first = next(i for i in range(n) if int(recs[i][nsq]) == k)
        order = [first] + [i for i in range(n) if i != first][:63]
        t = az.Board.GetEncodedStates([boards[i] for i in order], "cpu")
        out[f"planes_batch_k{k}"] = np.packbits(t.numpy().astype(bool), axis=None)
        out[f"planes_batch_k{k}_order"] = np.array(order, dtype=np.int32)
    mask = FourPlayerChess.get_legal_moves_mask(boards, "cpu")
    assert bool(((mask == 0) | (mask == 1)).all())
    out["mask"] = np.packbits(mask.numpy().astype(bool), axis=None)
    flat, flat_off, to_sq, results, heur = [], [0], [], [], []
    after_full, after_index = [], []
    for b, r in zip(boards, recs):
        lm = b.GetLegalMoves()
        keyed = sorted(lm, key=lambda m: m.GetFlatIndex())
        flat.extend(m.GetFlatIndex() for m in keyed)
        to_sq.extend(m.To().GetRow() * R + m.To().GetCol() for m in keyed)
        flat_off.append(len(flat))
        results.append(int(b.GetGameResult()))
        heur.append([b.CalculateHeuristic(az.RED_YELLOW), b.CalculateHeuristic(az.BLUE_GREEN)])
        if keyed:
            after_full.append(_record_of(b.TakeAction(keyed[-1]), R))
            after_index.append(_record_of(b.TakeAction(az.Move(keyed[-1].GetFlatIndex())), R))
        else:
            after_full.append(r)
            after_index.append(r)
    out["legal_flat"] = np.array(flat, dtype=np.int32)
    out["legal_to"] = np.array(to_sq, dtype=np.int32)
    out["legal_off"] = np.array(flat_off, dtype=np.int64)
    out["result"] = np.array(results, dtype=np.int32)
    out["heuristic"] = np.array(heur, dtype=np.int32)
    out["after_full"] = np.stack(after_full)
    out["after_index"] = np.stack(after_index)
    # (3) ParseActionspace as an index permutation per colour (src/cpp/board.cpp:257-263)
    ar = torch.arange(A * nsq, dtype=torch.float32).view(1, -1)
    out["parse_actionspace_perm"] = np.stack(
        [az.Board.ParseActionspace(ar, az.Player(az.PlayerColor(k))).contiguous().view(-1).numpy().astype(np.int32)
         for k in range(4)])
    out["statics"] = np.array([az.Board.num_state_channels, az.Board.state_space_size, A, az.Board.action_space_size,
                               az.Move.num_queen_moves_per_direction, az.Move.num_queen_moves,
                               az.Move.num_knight_moves, az.Board.invalidArea()], dtype=np.int64)
    path = os.path.join(HERE, f"binding_R{R}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))

    # (4) the reference's MCTS.search with the stand-in network (tests/golden/fake_net.py)
    mo = {}
    for case, (n_games, sims, start_ply) in {"a": (8, 40, 0), "b": (8, 64, 30)}.items():
        name = BINDING_NAMES[R][0]
        net = FakeNet(R)
        args = {"C": 3, "num_searches": sims, "pool_size": 1}
        if start_ply == 0:
            games = [FourPlayerChess(*parse_board_args_from_fen(getattr(start_fens, name).replace("\n", ""), R)) for _ in range(n_games)]
        else:  # mid-game positions of one mover colour (the reference steps its games in lock-step)
            turn0 = int(recs[start_ply // 3][nsq])
            idx = [i for i in range(start_ply // 3, n) if int(recs[i][nsq]) == turn0 and results[i] == 0][:n_games]
            games = [board_from_record(recs[i]) for i in idx]
        roots = MCTS(FourPlayerChess, net, args).search(games)
        mo[f"{case}_roots"] = np.stack([_record_of(g, R) for g in games])
        mo[f"{case}_sims"] = np.int64(sims)
        ch_off, ch_flat, ch_prior, ch_visits, ch_vsum, root_visits, n_nodes = [0], [], [], [], [], [], []

        def count(node):
            return 1 + sum(count(c) for c in node.GetChildren())

        for root in roots:
            for c in root.GetChildren():
                ch_flat.append(c.GetMoveMade().GetFlatIndex())
                ch_visits.append(c.GetVisitCount())
            ch_off.append(len(ch_flat))
            root_visits.append(root.GetVisitCount())
            n_nodes.append(count(root))
        mo[f"{case}_child_off"] = np.array(ch_off, dtype=np.int64)
        mo[f"{case}_child_flat"] = np.array(ch_flat, dtype=np.int32)
        mo[f"{case}_child_visits"] = np.array(ch_visits, dtype=np.int32)
        mo[f"{case}_root_visits"] = np.array(root_visits, dtype=np.int32)
        mo[f"{case}_n_nodes"] = np.array(n_nodes, dtype=np.int32)
        mo[f"{case}_nn_calls"] = np.int64(net.calls)
        mo[f"{case}_nn_positions"] = np.int64(net.positions)
    path = os.path.join(HERE, f"mcts_R{R}.npz")
    np.savez_compressed(path, **mo)
    print(path, os.path.getsize(path))


def make_viewer(R: int):
    """viewer_R<R>.npz: the pygame viewer's queries of the reference binding (src/cpp/board.cpp:50-57,120-232) on the
    positions of binding_R<R>.npz: per square the colours / teams that attack it (GetAttackedSquaresPlayers /
    GetAttackedSquaresTeams / IsAttackedByPlayer) and GetSimpleState's fields."""
    import alphazero_cpp as az

    nsq = R * R
    z = np.load(os.path.join(HERE, f"binding_R{R}.npz"))
    recs = z["recs"][:64]

    def board_from_record(rec):
        pieces = {}
        for sq in range(nsq):
            b = int(rec[sq])
            if b & 0x80:
                pieces[az.BoardLocation(sq // R, sq % R)] = az.Piece(az.PlayerColor((b >> 5) & 3), az.PieceType((b >> 2) & 7))
        return az.Board(az.Player(az.PlayerColor(int(rec[nsq]))), pieces)

    maps, simple_pieces, simple_turn = [], [], []
    for rec in recs:
        b = board_from_record(rec)
        m = np.zeros(nsq, dtype=np.uint8)
        players = b.GetAttackedSquaresPlayers()
        for color, locs in players.items():
            for loc in locs:
                m[loc.GetRow() * R + loc.GetCol()] |= 1 << int(color)
        for team, locs in b.GetAttackedSquaresTeams().items():
            for loc in locs:
                m[loc.GetRow() * R + loc.GetCol()] |= 16 << int(team)
        for sq in range(0, nsq, 7):  # the single-square query agrees with the map
            for c in range(4):
                assert b.IsAttackedByPlayer(az.BoardLocation(sq // R, sq % R), az.PlayerColor(c)) == bool((m[sq] >> c) & 1)
        st = b.GetSimpleState()
        assert {int(k): sorted(l.GetRow() * R + l.GetCol() for l in v) for k, v in st.attackedSquares.items()} == \
               {int(k): sorted(l.GetRow() * R + l.GetCol() for l in v) for k, v in players.items()}
        simple_turn.append(int(st.turn.GetColor()))
        simple_pieces.append(sorted((int(pp.GetPiece().GetColor()), int(pp.GetPiece().GetPieceType()),
                                     pp.GetLocation().GetRow() * R + pp.GetLocation().GetCol())
                                    for pl in st.pieces for pp in pl))
        assert len(st.castlingRights) == 4
        maps.append(m)
    flat = np.array([x for sp in simple_pieces for x in sp], dtype=np.int32).reshape(-1, 3)
    off = np.cumsum([0] + [len(sp) for sp in simple_pieces]).astype(np.int64)
    path = os.path.join(HERE, f"viewer_R{R}.npz")
    np.savez_compressed(path, recs=recs, attack_map=np.stack(maps), simple_turn=np.array(simple_turn, dtype=np.int32),
                        simple_pieces=flat, simple_pieces_off=off)
    print(path, os.path.getsize(path))


def main():
    if "--binding" in sys.argv:
        make_binding(int(sys.argv[sys.argv.index("--binding") + 1]))
        return
    if "--viewer" in sys.argv:
        make_viewer(int(sys.argv[sys.argv.index("--viewer") + 1]))
        return
    if "--viewer-only" in sys.argv:  # add the viewer fixtures without regenerating the others
        for R in (14, 8):
            bdir = os.path.join(ROOT, "oracle", "_ref", f"binding_R{R}")
            env = dict(os.environ, PYTHONPATH=os.pathsep.join([bdir, REF_PY, "/root/reference", ROOT]))
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--viewer", str(R)], env=env)
        return
    if not os.path.isdir(REF_PY):
        raise SystemExit("the reference tree is needed to regenerate the golden fixtures")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    make_engine()
    for R, IA in ((14, 3), (8, 2)):
        bdir = os.path.join(ROOT, "oracle", "_ref", f"binding_R{R}")
        if not os.path.exists(os.path.join(bdir, "alphazero_cpp.so")):
            subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref_binding.sh"), str(R), str(IA)])
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([bdir, REF_PY, "/root/reference", ROOT]))
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--binding", str(R)], env=env)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--viewer", str(R)], env=env)


if __name__ == "__main__":
    main()

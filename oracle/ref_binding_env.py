"""TEST / MEASUREMENT INFRASTRUCTURE.  The reference's own implementation of the WHOLE hot path --
rules engine AND tensor encoding -- timed on the host cores: the reference's pybind module
`alphazero_cpp` (built unmodified from /root/reference by oracle/build_ref_binding.sh into
oracle/_ref/binding_R<R>/) driven from Python exactly as `src/py/mcts.py` / `alphazero.py` drive it:

  per game and ply   state.GetGameResult()            engine/board.cpp:891-939
                     state.GetLegalMoves()            src/cpp/board.cpp:94-118
                     state.TakeAction(move)           src/cpp/board.cpp:234-239
  per batch and ply  Board.GetEncodedStates(states, "cpu")     src/cpp/board.cpp:305-356
                     the legal mask as FourPlayerChess.get_legal_moves_mask builds it
                     (src/py/four_player_chess_board.py:36-55): GetLegalMoves per state,
                     Board.GetLegalMovesIndices (src/cpp/board.cpp:424-449), four index tensors,
                     torch.zeros + index_put_.  (The bound C++ Board.GetLegalMovesMask, board.cpp:358-422,
                     raises a dtype error in index_put_ -- which is why the reference's Python does not use it.)

One Python process per host core (the reference holds the GIL for every call, wrapper.cpp has no
call_guard), each owning a slice of the resident games, torch intra-op threads = 1.  Games follow a
uniform random legal move (seeded); finished games are re-seeded from the start position.  Used only
by bench.py (`--impl reference`, `cpu_baseline`)."""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

_HERE = os.path.dirname(os.path.abspath(__file__))


def binding_dir(R: int) -> str:
    return os.path.join(_HERE, "_ref", f"binding_R{R}")


def available(R: int) -> bool:
    return os.path.exists(os.path.join(binding_dir(R), "alphazero_cpp.so"))


def _worker(R, start_rec, n_slots, steps, warmup, max_plies, seed, barrier, out_q, initial=None, initial_plies=None):
    try:
        _work(R, start_rec, n_slots, steps, warmup, max_plies, seed, barrier, out_q, initial, initial_plies)
    except BaseException as e:  # never leave the other processes waiting at a barrier
        barrier.abort()
        out_q.put(("error", repr(e)))


def _work(R, start_rec, n_slots, steps, warmup, max_plies, seed, barrier, out_q, initial=None, initial_plies=None):
    import random

    import torch
    torch.set_num_threads(1)
    sys.path.insert(0, binding_dir(R))
    import alphazero_cpp as az

    nsq = R * R
    A = az.Board.num_action_channels
    pieces = {}
    for sq in range(nsq):
        b = int(start_rec[sq])
        if b & 0x80:
            pieces[az.BoardLocation(sq // R, sq % R)] = az.Piece(az.PlayerColor((b >> 5) & 3),
                                                                az.PieceType((b >> 2) & 7))
    turn = az.Player(az.PlayerColor(int(start_rec[nsq])))

    def fresh():
        return az.Board(turn, pieces)

    def from_record(rec):
        pcs = {}
        for sq in range(nsq):
            b = int(rec[sq])
            if b & 0x80:
                pcs[az.BoardLocation(sq // R, sq % R)] = az.Piece(az.PlayerColor((b >> 5) & 3), az.PieceType((b >> 2) & 7))
        return az.Board(az.Player(az.PlayerColor(int(rec[nsq]))), pcs)

    rng = random.Random(seed)
    if initial is None:
        states = [fresh() for _ in range(n_slots)]
        plies = [0] * n_slots
    else:  # the slots start spread over whole games (positions and their ply numbers given by the caller)
        states = [from_record(initial[i]) for i in range(n_slots)]
        plies = [int(p) for p in initial_plies]
    one = torch.tensor(1, dtype=torch.float32)
    positions = 0
    t0 = 0.0
    for step in range(warmup + steps):
        if step == warmup:
            barrier.wait()
            t0 = time.perf_counter()
            positions = 0
        az.Board.GetEncodedStates(states, "cpu")
        legal = [s.GetLegalMoves() for s in states]
        b, pl, r, c = az.Board.GetLegalMovesIndices(legal, sum(len(m) for m in legal))
        mask = torch.zeros((n_slots, A, R, R), dtype=torch.float32)
        mask.index_put_((torch.tensor(b, dtype=torch.int64), torch.tensor(pl, dtype=torch.int64),
                         torch.tensor(r, dtype=torch.int64), torch.tensor(c, dtype=torch.int64)), one)
        for i in range(n_slots):
            s = states[i]
            done = s.GetGameResult() != az.GameResult.IN_PROGRESS
            if not done:
                moves = legal[i]
                if moves:
                    states[i] = s.TakeAction(moves[rng.randrange(len(moves))])
                    plies[i] += 1
                    done = plies[i] >= max_plies
                else:
                    done = True
            if done:
                states[i] = fresh()
                plies[i] = 0
            positions += 1
    dt = time.perf_counter() - t0
    barrier.wait()
    out_q.put((positions, dt))


def run(R, start_rec, n_games, steps, warmup, n_procs, max_plies=2048, seed=0x5EED, initial=None, initial_plies=None):
    """positions/s over all processes for `steps` plies of `n_games` resident games.  initial [n_games][REC] +
    initial_plies [n_games]: the positions the slots start from (default: every slot at the start record)."""
    ctx = mp.get_context("spawn")
    n_procs = max(1, min(n_procs, n_games))
    sizes = [n_games // n_procs + (1 if i < n_games % n_procs else 0) for i in range(n_procs)]
    barrier = ctx.Barrier(n_procs + 1)
    q = ctx.Queue()
    offs = [sum(sizes[:i]) for i in range(n_procs)]
    procs = [ctx.Process(target=_worker, args=(R, start_rec, sizes[i], steps, warmup, max_plies, seed + i, barrier, q,
                                               None if initial is None else initial[offs[i]: offs[i] + sizes[i]],
                                               None if initial is None else initial_plies[offs[i]: offs[i] + sizes[i]]))
             for i in range(n_procs)]
    for p in procs:
        p.start()
    try:
        barrier.wait(timeout=600)   # every worker has imported torch, built its boards and warmed up
        t0 = time.perf_counter()
        barrier.wait(timeout=3600)  # every worker has finished its timed steps
        dt = time.perf_counter() - t0
    except Exception:
        dt = float("nan")
    got = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    errs = [g[1] for g in got if g[0] == "error"]
    if errs:
        raise RuntimeError("reference binding worker failed: " + errs[0])
    positions = sum(g[0] for g in got)
    return {"positions_per_s": positions / dt, "positions": positions, "seconds": dt, "procs": n_procs}


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(_HERE))
    from alphazero_4_player_chess_b200.fen import start_record
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    print(run(14, start_record("STANDARD"), n, steps=4, warmup=1, n_procs=len(os.sched_getaffinity(0))))

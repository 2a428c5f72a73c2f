"""CPU, world_size 2 over gloo: the multi-GPU host logic (alphazero_4_player_chess_b200/shard.py).
Each rank plays its shard of the playout games with the CPU oracle standing in for the kernels; the
gathered results must equal a single-process run of all games -- the partition is disjoint, complete
and invariant to the number of ranks -- and the counter reduction must sum over ranks."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.shard import Shard, max_over_ranks, reduce_counters
from tests.util import SEED, oracle_for

SLOTS, GENERATIONS, MAX_PLIES = 3, 2, 30


def play_shard(shard: Shard):
    """[(game id, positions, checksum of the final board)] + counters [positions, finished]."""
    o = oracle_for(8)
    start = start_record("EIGHT_SIMPLE")
    rows, positions = [], 0
    for slot in range(shard.games_per_gpu):
        for gen in range(GENERATIONS):
            gid = shard.game_id(slot, gen)
            p = o.playout(start, SEED, gid, MAX_PLIES)
            rows.append((gid, p["n"], int(p["recs"][-1].astype(np.int64).sum())))
            positions += p["n"]
    return rows, torch.tensor([positions, len(rows)], dtype=torch.int64)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = Shard(rank, world, SLOTS)
    rows, counters = play_shard(shard)
    total = reduce_counters(counters)
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    slowest = max_over_ranks(float(rank + 1))
    if rank == 0:
        q.put((sorted(r for part in gathered for r in part), total.tolist(), slowest))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows2, total2, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the same games on one rank with twice the slots: ids 0..5 (generation 0) and 6..11 (generation 1)
    rows1, total1 = play_shard(Shard(0, 1, 2 * SLOTS))
    assert rows2 == sorted(rows1)
    assert total2 == total1.tolist()
    assert slowest == 2.0
    ids = [r[0] for r in rows2]
    assert ids == list(range(2 * SLOTS * GENERATIONS))  # disjoint and complete


def test_owner_is_the_inverse_of_game_id():
    for world in (1, 2, 4, 8):
        for rank in range(world):
            sh = Shard(rank, world, 4096)
            for slot in (0, 1, 4095):
                for gen in (0, 1, 7):
                    assert sh.owner(sh.game_id(slot, gen)) == (rank, slot, gen)

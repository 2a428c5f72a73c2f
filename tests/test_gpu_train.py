"""-m gpu: one iteration of the outer loop (train.py: self-play -> replay -> optimiser steps -> validation)
with a tiny network at 8x8: the loop runs end to end on the device and the optimiser changes the weights."""
import pytest
import torch

from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.net import AutocastNet, PolicyValueNet
from alphazero_4_player_chess_b200.selfplay import SelfPlay
from alphazero_4_player_chess_b200.train import Learner

pytestmark = pytest.mark.gpu


def test_one_learning_iteration():
    torch.manual_seed(0)
    R = 8
    model = PolicyValueNet(R, blocks=1, hidden=16, device="cuda")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    args = {"C": 3, "num_searches": 8, "temperature": 1.1, "max_game_length": 10, "heuristic_weight": 0.02,
            "batch_size": 32, "replay_buffer_capacity": 4096, "validation_buffer_capacity": 1024,
            "num_iterations": 1, "num_games": 32, "num_parallel_games": 32}
    sp = SelfPlay(R, 32, AutocastNet(model), args, start_record("EIGHT_SIMPLE"))
    learner = Learner(sp, model, opt, args)
    before = [p.detach().clone() for p in model.parameters()]
    out = learner.learn(1)
    assert out[0]["train_steps"] >= 1 and out[0]["replay"] >= 32
    assert all(torch.isfinite(torch.tensor([e["policy_loss"], e["value_loss"]])).all() for e in learner.log)
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    assert out[0]["validation"] is None or out[0]["validation"]["validation_loss"] > 0
    boards, flat, visits, value = learner.train_buf.sample(8)
    assert boards.shape == (8, 80) and bool((visits.sum(dim=1) > 0).all()) and bool((value.abs() <= 1.0).all())


def _ddp_worker(rank, world, port, q):
    import os

    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    dev = f"cuda:{rank}"
    R = 8
    torch.manual_seed(0)  # identical initial weights
    model = PolicyValueNet(R, blocks=1, hidden=16, device=dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    args = {"C": 3, "num_searches": 6, "temperature": 1.1, "max_game_length": 8 + 4 * rank, "heuristic_weight": 0.02,
            "batch_size": 32, "replay_buffer_capacity": 4096, "validation_buffer_capacity": 1024,
            "num_iterations": 1, "num_games": 32, "num_parallel_games": 32}
    torch.manual_seed(100 + rank)  # every rank plays its own games, of its own length: different replay sizes
    sp = SelfPlay(R, 32, AutocastNet(model), args, start_record("EIGHT_SIMPLE"), device=dev)
    learner = Learner(sp, model, opt, args)
    assert learner.ddp is not None
    out = learner.learn(1)
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(learner.train_buf)], dtype=torch.int64, device=dev))
    if rank == 0:
        q.put((out[0]["train_steps"], bool(all(torch.equal(gathered[0], g) for g in gathered[1:])),
               [int(s) for s in sizes]))
    dist.barrier()
    dist.destroy_process_group()


def test_learner_over_two_gpus_nccl():
    """f4 on real GPUs: one process per GPU, each with its own self-play shard (games of different lengths, so the
    replay sizes differ), DistributedDataParallel over NCCL: the ranks run the same number of optimiser steps and end
    with identical weights."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket

    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    steps, same, sizes = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert steps >= 1 and same, (steps, same, sizes)
    assert sizes[0] != sizes[1], sizes

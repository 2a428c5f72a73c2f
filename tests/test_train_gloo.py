"""CPU, world_size 2 over gloo: the multi-GPU side of the learner (train.py) -- every rank feeds its own replay
shard, DistributedDataParallel averages the gradients, so the ranks' weights stay identical after a step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphazero_4_player_chess_b200.geometry import GEOMETRIES
from alphazero_4_player_chess_b200.train import Learner


class TinyNet(torch.nn.Module):
    def __init__(self, g):
        super().__init__()
        self.p = torch.nn.Linear(g.state_space_size, g.action_space_size)
        self.v = torch.nn.Linear(g.state_space_size, 1)

    def forward(self, x):
        x = x.flatten(1)
        return self.p(x), torch.tanh(self.v(x))


class StubSelfPlay:
    """Stands in for SelfPlay on a CPU box: same attributes the Learner touches, synthetic replay content."""

    def __init__(self, R, rank):
        self.geom = GEOMETRIES[R]
        self.device = torch.device("cpu")
        self.hist_flat = torch.zeros((1, 1, 300), dtype=torch.int32)
        self.gen = torch.Generator().manual_seed(100 + rank)

    def replay(self, m):
        g = self.geom
        flat = torch.randint(0, g.action_space_size, (m, 300), generator=self.gen, dtype=torch.int32)
        visits = torch.randint(1, 5, (m, 300), generator=self.gen, dtype=torch.int32)
        return {"boards": torch.randint(0, 255, (m, g.record_bytes), generator=self.gen, dtype=torch.uint8),
                "child_flat": flat, "child_visits": visits, "value": torch.rand(m, generator=self.gen) * 2 - 1}

    def encoded_states(self, rp):
        return (rp["boards"][:, :1].float() / 255.0).expand(-1, self.geom.state_space_size).reshape(-1, 24, self.geom.R, self.geom.R).contiguous()

    def policy_targets(self, rp):
        p = torch.zeros((rp["child_flat"].shape[0], self.geom.action_space_size))
        p.scatter_add_(1, rp["child_flat"].long(), rp["child_visits"].float())
        return p / p.sum(dim=1, keepdim=True)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical initial weights on every rank
    sp = StubSelfPlay(8, rank)
    model = TinyNet(sp.geom)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    args = {"batch_size": 16, "replay_buffer_capacity": 256, "validation_buffer_capacity": 64}
    learner = Learner(sp, model, opt, args)
    assert learner.ddp is not None
    learner.store(sp.replay(200 if rank == 0 else 90))  # different data AND a different amount of it on every rank
    local_steps = -(-len(learner.train_buf) // 16)
    steps = learner.train()  # must not hang: the ranks agree on the smaller step count first
    counts = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_steps, steps]))
    assert counts[0][1] == counts[1][1] == min(int(counts[0][0]), int(counts[1][0])), counts
    assert int(counts[0][0]) != int(counts[1][0])
    # a rank with less than one batch keeps everybody from training (no collective is left half-entered)
    starved = Learner(StubSelfPlay(8, rank), model, opt, args)
    starved.ddp = learner.ddp
    starved.store(sp.replay(200 if rank == 0 else 4))
    assert starved.train() == 0
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put((steps, bool(torch.equal(gathered[0], gathered[1])), float(flat.abs().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_keeps_the_ranks_in_step():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    steps, same, norm = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert steps >= 1 and same and norm > 0

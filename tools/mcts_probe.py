"""Runs a short batched PUCT search with fixed network outputs (tree kernels only); for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.mcts import BatchedMCTS

n, sims, R = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 40, 14
m = BatchedMCTS(R, n, None, {"C": 3, "num_searches": sims})
roots = torch.from_numpy(start_record("STANDARD")).unsqueeze(0).repeat(n, 1)
torch.manual_seed(0)
logits = torch.randn((n, m.geom.action_space_size), device="cuda")
values = torch.tanh(torch.randn(n, device="cuda"))
m.reset(roots)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(sims):
    if i == sims // 2:
        e0.record()
    m.select()
    m.expand_backup(logits, values)
e1.record()
torch.cuda.synchronize()
m.check_errors()
print(f"{(sims - sims // 2)} sim batches of {n} games: {e0.elapsed_time(e1) / (sims - sims // 2) * 1e3:.1f} us per batch; nodes {int(m.n_nodes.sum())}")

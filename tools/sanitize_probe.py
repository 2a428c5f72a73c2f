"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck): observe with all outputs,
playout steps (sync and async dense), make full / index, heuristic, encode, a short PUCT search, perft(3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.mcts import BatchedMCTS
from alphazero_4_player_chess_b200.perft import perft

for name, R, n in (("STANDARD", 14, 67), ("THIRTEEN", 13, 33), ("EIGHT_SIMPLE", 8, 45)):
    start = start_record(name, castling=True)
    env = BatchedEnv(R, n)
    env.reset_playout(start)
    for step in range(60):
        env.playout_step(max_plies=50, planes=True, mask=True, chosen=True, async_dense=bool(step % 2))
    env.join()
    env.observe(planes=True, mask=True, moves=True, flat=True, k=2)
    env.encode(k=torch.arange(n, dtype=torch.int32, device="cuda") % 4)
    mv = env.moves_buffer()[:, 0].contiguous()
    env.make_moves(mv)
    env.observe(planes=False, mask=False, flat=True)
    env.make_index(env.flat_buffer()[:, 0].contiguous())
    env.heuristic()
    torch.manual_seed(0)
    m = BatchedMCTS(R, n, None, {"C": 3, "num_searches": 12})
    m.reset(torch.from_numpy(start).unsqueeze(0).repeat(n, 1))
    logits = torch.randn((n, m.geom.action_space_size), device="cuda")
    values = torch.tanh(torch.randn(n, device="cuda"))
    for _ in range(12):
        m.select()
        m.expand_backup(logits, values)
    torch.cuda.synchronize()
    m.check_errors()
    print(name, "perft", perft(R, start, 3), "nodes", int(m.n_nodes.sum()))
print("sanitize probe done")

"""A few dense playout steps (rules_kernel + expand_kernel) from a fast-forwarded position mix; for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import start_record
ff = int(sys.argv[1]) if len(sys.argv) > 1 else 200
env = BatchedEnv(14, 4096)
env.reset_playout(start_record("STANDARD", castling=True))
for _ in range(ff):
    env.playout_step(planes=False, mask=False)
for _ in range(6):
    env.playout_step(planes=True, mask=True, async_dense=True)
env.join()
torch.cuda.synchronize()
print("ok", float(env.counters[6]) / float(env.counters[0]))

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.mcts import BatchedMCTS
from alphazero_4_player_chess_b200.net import InferenceNet, PolicyValueNet
env = BatchedEnv(R=14, n_games=4096)
env.load(start_record("STANDARD"))
env.observe(planes=True, mask=True, flat=True)
planes, mask = env.planes_buffer(), env.mask_buffer()
net = InferenceNet(PolicyValueNet(14, blocks=2, hidden=32))
search = BatchedMCTS(14, 1024, net, {"C": 3, "num_searches": 8})
search.search(env.boards[:1024])
probs = search.action_probs()
print(planes.shape, mask.shape, probs.shape, float(probs.sum()), int(env.counts[0]))

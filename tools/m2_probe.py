"""configs[3] probe: 1,024 games x N sims of batched PUCT with the random-init ResNet 10x128, fp32 and bf16."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.mcts import BatchedMCTS
from alphazero_4_player_chess_b200.net import InferenceNet, PolicyValueNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 400
R = 14
roots = torch.from_numpy(start_record("STANDARD")).unsqueeze(0).repeat(n, 1)
for bf16 in (True, False):
    torch.manual_seed(0)
    net = InferenceNet(PolicyValueNet(R, 10, 128, device="cuda"), bf16=bf16)
    m = BatchedMCTS(R, n, net, {"C": 3, "num_searches": sims})
    m.args["num_searches"] = 4
    m.search(roots)
    torch.cuda.synchronize()
    m.args["num_searches"] = sims
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.search(roots)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    m.check_errors()
    # network alone
    x = m.planes
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        net(x)
    e1.record()
    torch.cuda.synchronize()
    net_ms = e0.elapsed_time(e1) / 20
    print(f"{'bf16' if bf16 else 'fp32'}: {n} games x {sims} sims: {ms:.0f} ms, {n * sims / ms * 1e3:.0f} sims/s, "
          f"{ms / sims:.2f} ms/sim-batch, net alone {net_ms:.2f} ms, nodes {int(m.n_nodes.sum())}, "
          f"max nodes/game {int(m.n_nodes.max())} of {m.node_cap}, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del m, net
    torch.cuda.empty_cache()

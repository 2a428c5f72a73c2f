"""Times the reference's MCTS call sequence (tests/test_gpu_dropin.py::drive_search) through a pybind module
named alphazero_cpp: ours (dropin/) or the unmodified reference's (oracle/_ref/binding_R14).
usage: dropin_probe.py ours|ref [games] [sims]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
which = sys.argv[1]
games_n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 50
R = 14
import torch
if which == "ours":
    from alphazero_4_player_chess_b200 import build
    build.build_binding()
    sys.path.insert(0, build.DROPIN)
else:
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "binding_R14"))
import alphazero_cpp as az
if which == "ours":
    az.set_board_size(R)
from alphazero_4_player_chess_b200.fen import start_record
from tests.golden.fake_net import FakeNet
from tests.test_gpu_dropin import board_from_record, drive_search

rec = start_record("STANDARD")
net = FakeNet(R, device="cuda")
games = [board_from_record(az, rec, R) for _ in range(8)]
drive_search(az, games, net, 3, 4)  # warm-up
torch.cuda.synchronize()
reps = 1 if "--profile" in sys.argv else 3
rates = []
for rep in range(reps):
    games = [board_from_record(az, rec, R) for _ in range(games_n)]
    if "--profile" in sys.argv:
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
    t0 = time.perf_counter()
    roots = drive_search(az, games, net, 3, sims)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if "--profile" in sys.argv:
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    rates.append(games_n * sims / dt)
    visits = sum(r.GetVisitCount() for r in roots)
    del roots, games
rates.sort()
print(f"{which}: {games_n} games x {sims} sims, median of {reps} = {rates[len(rates) // 2]:.0f} sims/s (all: "
      f"{', '.join(f'{r:.0f}' for r in rates)}; root visits {visits})", flush=True)

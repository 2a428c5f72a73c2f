"""Ad-hoc timing probe for the dense-output pipeline (not part of the product)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_4_player_chess_b200 import _lib
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import start_record

L = _lib.lib()
env = BatchedEnv(14, 4096)
env.reset_playout(start_record("STANDARD", castling=True))
for _ in range(50):
    env.playout_step(planes=False, mask=False)

def timeit(name, fn, reps=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / reps * 1e3:8.1f} us/call")

timeit("playout rules only", lambda: env.playout_step(planes=False, mask=False))
timeit("playout planes only", lambda: env.playout_step(planes=True, mask=False))
timeit("playout mask only", lambda: env.playout_step(planes=False, mask=True))
timeit("playout planes+mask", lambda: env.playout_step(planes=True, mask=True))
timeit("playout planes+mask async", lambda: env.playout_step(planes=True, mask=True, async_dense=True))
env.join()
timeit("playout planes+mask incremental", lambda: env.playout_step(planes=True, mask=True, incremental=True))
timeit("encode only", lambda: env.encode())
timeit("observe planes+mask (no playout)", lambda: env.observe(planes=True, mask=True))
timeit("observe mask", lambda: env.observe(planes=False, mask=True))
timeit("torch zero_ mask", lambda: env.mask_buffer().zero_())

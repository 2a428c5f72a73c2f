"""Diagnostics: the dense playout step (rules_kernel + expand_kernel) under the launch knobs of the FPC_EXPERIMENT
build (tools/build_experiment.sh -> tools/libfpc_x.so).  usage: FPC_LIB_PATH=tools/libfpc_x.so FPC_X_...=1 python
tools/overlap_probe.py [steps] [fast_forward_plies]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphazero_4_player_chess_b200 import _lib
from alphazero_4_player_chess_b200.env import BatchedEnv
from alphazero_4_player_chess_b200.fen import start_record

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
ff = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = _lib.lib()
env = BatchedEnv(14, 4096)
env.reset_playout(start_record("STANDARD", castling=True))
for _ in range(ff):
    env.playout_step(planes=False, mask=False)
for _ in range(20):
    env.playout_step(planes=True, mask=True, async_dense=True)
env.join()
torch.cuda.synchronize()
_lib.check(L.fpc_profile_enable(1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
env.counters.zero_()
e0.record()
for _ in range(steps):
    env.playout_step(planes=True, mask=True, async_dense=True)
env.join()
e1.record()
torch.cuda.synchronize()
n, ex, ru = ctypes.c_int(0), ctypes.c_double(0), ctypes.c_double(0)
_lib.check(L.fpc_profile_read(ctypes.byref(n), ctypes.byref(ex), ctypes.byref(ru)))
r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
r0.record()
for _ in range(steps):
    env.playout_step(planes=False, mask=False)
r1.record()
torch.cuda.synchronize()
knobs = {k: v for k, v in os.environ.items() if k.startswith("FPC_X_")}
c = env.counters.cpu()
print(f"{knobs}: step {e0.elapsed_time(e1) / steps * 1e3:.1f} us, expand in-loop {ex.value / max(n.value, 1) * 1e3:.1f} us, "
      f"rules in-loop {ru.value / max(n.value, 1) * 1e3:.1f} us, rules-only step {r0.elapsed_time(r1) / steps * 1e3:.1f} us, "
      f"avg legal {float(c[6]) / max(float(c[0]), 1):.1f}", flush=True)

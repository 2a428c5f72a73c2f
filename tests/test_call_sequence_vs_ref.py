"""CPU, build container only (needs /root/reference and oracle/_ref/binding_R*): the call sequence the GPU drop-in tests
and the drop-in benchmark drive (`tests/test_gpu_dropin.py::drive_search`) IS the reference's: on the reference's own
binding it builds the same trees as the reference's real `MCTS.search` (`/root/reference/src/py/mcts.py`, imported from
where it lies) and as the committed fixtures that search produced."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = "/root/reference/src/py"

_CHILD = r'''
import sys, os
import numpy as np, torch
R = int(sys.argv[1])
import alphazero_cpp as az          # the UNMODIFIED reference binding (oracle/_ref/binding_R<R>)
from mcts import MCTS               # the reference's own search, /root/reference/src/py/mcts.py
from four_player_chess_board import FourPlayerChess
from tests.golden.fake_net import FakeNet
from tests.test_gpu_dropin import drive_search
import tests.test_gpu_dropin as T

assert az.Board.nRows() == R and not hasattr(az, "set_board_size")
nsq = R * R
z = np.load(os.path.join("tests", "golden", f"mcts_R{R}.npz"))

def board(rec, cls):
    pieces = {}
    for sq in range(nsq):
        b = int(rec[sq])
        if b & 0x80:
            pieces[az.BoardLocation(sq // R, sq % R)] = az.Piece(az.PlayerColor((b >> 5) & 3), az.PieceType((b >> 2) & 7))
    return cls(az.Player(az.PlayerColor(int(rec[nsq]))), pieces)

# the emulation builds its masks on "cuda"; on this box everything is CPU
T.legal_moves_mask.__defaults__ = None
_orig = T.legal_moves_mask
T.legal_moves_mask = lambda az_, states, device: _orig(az_, states, "cpu")
class CpuBoard:
    """az.Board with device strings forced to "cpu" (no GPU here)."""
    def __getattr__(self, name):
        return getattr(az.Board, name)
    def GetEncodedStates(self, states, device):
        return az.Board.GetEncodedStates(states, "cpu")
class AzCpu:
    def __getattr__(self, name):
        return CpuBoard() if name == "Board" else getattr(az, name)

for case in ("a", "b"):
    roots_rec, sims = z[f"{case}_roots"], int(z[f"{case}_sims"])
    ref_roots = MCTS(FourPlayerChess, FakeNet(R), {"C": 3, "num_searches": sims, "pool_size": 1}).search(
        [board(r, FourPlayerChess) for r in roots_rec])
    emu_roots = drive_search(AzCpu(), [board(r, az.Board) for r in roots_rec], FakeNet(R), 3, sims)
    off = z[f"{case}_child_off"]
    for g, (a, b) in enumerate(zip(ref_roots, emu_roots)):
        ca, cb = a.GetChildren(), b.GetChildren()
        fa = [c.GetMoveMade().GetFlatIndex() for c in ca]
        assert fa == [c.GetMoveMade().GetFlatIndex() for c in cb], (case, g)
        va = [c.GetVisitCount() for c in ca]
        assert va == [c.GetVisitCount() for c in cb], (case, g)
        assert a.GetVisitCount() == b.GetVisitCount() == int(z[f"{case}_root_visits"][g])
        assert fa == z[f"{case}_child_flat"][off[g]: off[g + 1]].tolist() and va == z[f"{case}_child_visits"][off[g]: off[g + 1]].tolist()
        # one level down as well
        for x, y in zip(ca, cb):
            assert [c.GetVisitCount() for c in x.GetChildren()] == [c.GetVisitCount() for c in y.GetChildren()]
print("ok")
'''


@pytest.mark.parametrize("R", [8, 14])
def test_emulated_call_sequence_equals_the_reference_search(R):
    bdir = os.path.join(ROOT, "oracle", "_ref", f"binding_R{R}")
    if not (os.path.isdir(REF_PY) and os.path.exists(os.path.join(bdir, "alphazero_cpp.so"))):
        pytest.skip("needs the reference tree and its binding (build container only)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([bdir, REF_PY, "/root/reference", ROOT]), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", _CHILD, str(R)], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-4000:]

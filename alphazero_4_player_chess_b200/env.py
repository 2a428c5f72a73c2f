"""BatchedEnv: N concurrent four-player-chess games resident in HBM.

Host-side mirror of the reference's `alphazero_cpp.Board` batch calls (GetLegalMoves,
TakeAction, GetGameResult, GetEncodedStates, get_legal_moves_mask; `src/cpp/wrapper.cpp:165-226`,
`src/py/four_player_chess_board.py:36-55`) on top of the C-ABI in include/fpc.h.  PyTorch is used
for device memory and streams only; every rule runs in libfpc.so's CUDA kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import FPC_MAX_MOVES, check
from .geometry import GEOMETRIES, NUM_STATE_CHANNELS, Geometry


def _ptr(t):
    return None if t is None else t.data_ptr()


class BatchedEnv:
    def __init__(self, R: int = 14, n_games: int = 4096, device: str | torch.device = "cuda"):
        self.geom: Geometry = GEOMETRIES[R]
        self.R = R
        self.n = int(n_games)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.FpcError("BatchedEnv needs a CUDA device (there is no CPU fallback)")
        self.L = _lib.lib()
        rec = self.geom.record_bytes
        self.boards = torch.zeros((self.n, rec), dtype=torch.uint8, device=self.device)
        self.counts = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        # playout state
        self.game = torch.arange(self.n, dtype=torch.int64, device=self.device)
        self.ply = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        self.start = torch.zeros(rec, dtype=torch.uint8, device=self.device)
        self.counters = torch.zeros(8, dtype=torch.int64, device=self.device)
        self.chosen = torch.zeros(self.n, dtype=torch.int64, device=self.device)
        self._planes = None
        self._mask = None
        # the handle that lets FPC_FLAG_INCREMENTAL update this env's own planes / mask buffers in place; it lives and
        # dies with them (nothing is inferred from tensor addresses)
        self._track = None
        self._moves = None
        self._flat = None

    # ---- buffers -----------------------------------------------------------------------------
    def planes_buffer(self) -> torch.Tensor:
        if self._planes is None:
            self._planes = torch.empty((self.n, NUM_STATE_CHANNELS, self.R, self.R), dtype=torch.float32,
                                       device=self.device)
        return self._planes

    def mask_buffer(self) -> torch.Tensor:
        if self._mask is None:
            self._mask = torch.empty((self.n, self.geom.num_action_channels, self.R, self.R),
                                     dtype=torch.float32, device=self.device)
        return self._mask

    def moves_buffer(self) -> torch.Tensor:
        if self._moves is None:
            self._moves = torch.zeros((self.n, FPC_MAX_MOVES), dtype=torch.int64, device=self.device)
        return self._moves

    def flat_buffer(self) -> torch.Tensor:
        if self._flat is None:
            self._flat = torch.zeros((self.n, FPC_MAX_MOVES), dtype=torch.int32, device=self.device)
        return self._flat

    # ---- zero-copy hand-off (DLPack) --------------------------------------------------------------
    def dlpack(self, name: str = "planes"):
        """DLPack capsule of a device buffer ("planes", "mask", "boards", "counts", "status", "moves", "flat"):
        the same memory the kernels write, for any DLPack consumer (`torch.from_dlpack`, CuPy, JAX ...)."""
        from torch.utils.dlpack import to_dlpack
        buf = {"planes": self.planes_buffer, "mask": self.mask_buffer, "moves": self.moves_buffer,
               "flat": self.flat_buffer, "boards": lambda: self.boards, "counts": lambda: self.counts,
               "status": lambda: self.status}[name]()
        return to_dlpack(buf)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def track(self):
        if self._track is None:
            with torch.cuda.device(self.device):
                self._track = self.L.fpc_dense_track_create(self.R, self.n)
            if not self._track:
                raise _lib.FpcError(self.L.fpc_last_error().decode())
        return self._track

    def invalidate_dense(self) -> None:
        """Declare that something else wrote to planes_buffer() / mask_buffer(): the next incremental call rewrites them."""
        if self._track is not None:
            self.L.fpc_dense_track_invalidate(self._track)

    def close(self) -> None:
        if getattr(self, "_track", None):
            self.L.fpc_dense_track_destroy(self._track)
            self._track = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state -------------------------------------------------------------------------------
    def load(self, records: np.ndarray | torch.Tensor) -> None:
        """Load board records: one [REC] record (broadcast to all games) or [n][REC]."""
        t = torch.as_tensor(np.ascontiguousarray(records) if isinstance(records, np.ndarray) else records,
                            dtype=torch.uint8)
        if t.dim() == 1:
            t = t.unsqueeze(0).expand(self.n, -1)
        self.boards.copy_(t.to(self.device))

    def set_start(self, record: np.ndarray) -> None:
        self.start.copy_(torch.as_tensor(np.ascontiguousarray(record), dtype=torch.uint8).to(self.device))

    def reset_playout(self, record: np.ndarray, first_game: int = 0) -> None:
        self.set_start(record)
        self.load(record)
        self.game.copy_(torch.arange(first_game, first_game + self.n, dtype=torch.int64))
        self.ply.zero_()
        self.counters.zero_()

    # ---- kernels -----------------------------------------------------------------------------
    def join(self) -> None:
        """Order the current stream after the latest dense expansion (after async_dense calls)."""
        with torch.cuda.device(self.device):
            check(self.L.fpc_join(self._stream()))

    def observe(self, planes: bool = True, mask: bool = True, moves: bool = False, flat: bool = False,
                k: int | torch.Tensor = -1, async_dense: bool = False, incremental: bool = False):
        """Legal moves / result / planes / mask of every game: rules_kernel, plus expand_kernel for the dense tensors
        (async_dense: do not wait for it, see join(); incremental: update the resident tensors in place)."""
        with torch.cuda.device(self.device):
            d_k = k if isinstance(k, torch.Tensor) else None
            check(self.L.fpc_observe_tracked(
                self.track() if (planes or mask) else None, self.R, self.boards.data_ptr(), self.n,
                _ptr(self.moves_buffer() if moves else None), _ptr(self.flat_buffer() if flat else None),
                self.counts.data_ptr(), self.status.data_ptr(),
                _ptr(self.planes_buffer() if planes else None), _ptr(d_k), -1 if d_k is not None else int(k),
                _ptr(self.mask_buffer() if mask else None), int(async_dense) | (2 if incremental else 0),
                self._stream()))
        return self

    def encode(self, k: int | torch.Tensor = -1) -> torch.Tensor:
        with torch.cuda.device(self.device):
            d_k = k if isinstance(k, torch.Tensor) else None
            check(self.L.fpc_observe_tracked(self.track(), self.R, self.boards.data_ptr(), self.n, None, None, None, None,
                                             self.planes_buffer().data_ptr(), _ptr(d_k),
                                             -1 if d_k is not None else int(k), None, 0, self._stream()))
        return self._planes

    def make_moves(self, moves: torch.Tensor) -> torch.Tensor:
        err = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.fpc_make_moves(self.R, self.boards.data_ptr(), moves.data_ptr(), self.n,
                                        self.boards.data_ptr(), err.data_ptr(), self._stream()))
        return err

    def make_index(self, flat: torch.Tensor) -> torch.Tensor:
        err = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.fpc_make_index(self.R, self.boards.data_ptr(), flat.data_ptr(), self.n,
                                        self.boards.data_ptr(), err.data_ptr(), self._stream()))
        return err

    def heuristic(self) -> torch.Tensor:
        out = torch.zeros(self.n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.L.fpc_heuristic(self.R, self.boards.data_ptr(), self.n, out.data_ptr(), self._stream()))
        return out

    def playout_stepper(self, seed: int = 0x5EED, max_plies: int = 2048, game_stride: int | None = None,
                        planes: bool = True, mask: bool = True, k: int = -1, async_dense: bool = False,
                        incremental: bool = False):
        """playout_step with every argument resolved once: returns a zero-argument callable that issues the same
        fpc_playout_step_tracked call (a few microseconds of host time instead of the ~30 us the tensor look-ups of
        playout_step cost per call).  The env's device must be the current device and the stream current now must
        stay the one the caller works on; buffers are the env's own."""
        if torch.cuda.current_device() != (self.device.index if self.device.index is not None else torch.cuda.current_device()):
            raise _lib.FpcError("playout_stepper: make the env's device current first (torch.cuda.set_device)")
        import ctypes as C
        vp = C.c_void_p
        stride = self.n if game_stride is None else game_stride
        args = (vp(self.track() if (planes or mask) else None), C.c_int(self.R), vp(self.boards.data_ptr()), C.c_int(self.n),
                C.c_uint64(seed), vp(self.game.data_ptr()), vp(self.ply.data_ptr()), vp(self.start.data_ptr()),
                C.c_int(max_plies), C.c_uint64(stride), vp(None), vp(self.counts.data_ptr()), vp(self.status.data_ptr()),
                vp(_ptr(self.planes_buffer() if planes else None)), vp(None), C.c_int(int(k)),
                vp(_ptr(self.mask_buffer() if mask else None)), vp(self.counters.data_ptr()),
                C.c_int(int(async_dense) | (2 if incremental else 0)), vp(self._stream()))
        fn = self.L.fpc_playout_step_tracked

        def step() -> None:
            rc = fn(*args)
            if rc:
                check(rc)
        return step

    def playout_step(self, seed: int = 0x5EED, max_plies: int = 2048, game_stride: int | None = None,
                     planes: bool = True, mask: bool = True, k: int = -1, chosen: bool = False,
                     async_dense: bool = False, incremental: bool = False) -> None:
        """One ply for every game slot (BASELINE.json configs[1]); finished slots are re-seeded."""
        stride = self.n if game_stride is None else game_stride
        with torch.cuda.device(self.device):
            check(self.L.fpc_playout_step_tracked(
                self.track() if (planes or mask) else None, self.R, self.boards.data_ptr(), self.n, seed, self.game.data_ptr(), self.ply.data_ptr(),
                self.start.data_ptr(), max_plies, stride, _ptr(self.chosen if chosen else None),
                self.counts.data_ptr(), self.status.data_ptr(),
                _ptr(self.planes_buffer() if planes else None), None, int(k),
                _ptr(self.mask_buffer() if mask else None), self.counters.data_ptr(),
                int(async_dense) | (2 if incremental else 0), self._stream()))

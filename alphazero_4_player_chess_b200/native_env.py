"""NativeEnv: the environment-owned board store of the C-ABI (`fpc_env_*`, include/fpc.h) -- device memory, stream and
DLPack export all inside libfpc.so, nothing from PyTorch.  This is the hand-off for consumers that are not PyTorch
(boards and encoded planes go to the network zero-copy via DLPack): `dlpack("planes")` returns a PyCapsule
named "dltensor" over a DLManagedTensor the library built itself; `torch.from_dlpack`, CuPy or JAX wrap it without a
copy, and the store stays alive until the last such tensor is released.

BatchedEnv (env.py) is the PyTorch-owned twin: same kernels, tensors allocated by torch."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .geometry import GEOMETRIES

_NAMES = {"boards": _lib.ENV_BOARDS, "counts": _lib.ENV_COUNTS, "status": _lib.ENV_STATUS, "moves": _lib.ENV_MOVES,
          "flat": _lib.ENV_FLAT, "planes": _lib.ENV_PLANES, "mask": _lib.ENV_MASK, "ply": _lib.ENV_PLY,
          "game": _lib.ENV_GAME}
_CAPSULE_NAME = b"dltensor"  # must outlive every capsule

_new_capsule = C.pythonapi.PyCapsule_New
_new_capsule.restype = C.py_object
_new_capsule.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]


class NativeEnv:
    def __init__(self, R: int = 14, n_games: int = 4096, device: int = 0):
        self.geom = GEOMETRIES[R]
        self.R, self.n, self.device = R, int(n_games), int(device)
        self.L = _lib.lib()
        self._h = self.L.fpc_env_create(self.device, R, self.n)
        if not self._h:
            raise _lib.FpcError(self.L.fpc_last_error().decode())

    def _which(self, outputs) -> int:
        w = 0
        for name in outputs:
            w |= _NAMES[name]
        return w

    def load(self, records: np.ndarray, first: int = 0) -> None:
        rec = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, self.geom.record_bytes)
        check(self.L.fpc_env_set_boards(self._h, rec.ctypes.data, first, len(rec)))

    def boards(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.empty((count, self.geom.record_bytes), dtype=np.uint8)
        check(self.L.fpc_env_get_boards(self._h, out.ctypes.data, first, count))
        return out

    def observe(self, outputs=("planes", "mask"), k: int = -1) -> None:
        """fpc_observe over the store (asynchronous on the environment's stream): counts and status always, plus the
        named outputs.  k = -1 rotates every game by its own side to move."""
        check(self.L.fpc_env_observe(self._h, self._which(outputs), k))

    def playout_step(self, start: np.ndarray, seed: int, max_plies: int = 2048, outputs=(), k: int = -1,
                     game_stride: int | None = None) -> None:
        s = np.ascontiguousarray(start, dtype=np.uint8)
        check(self.L.fpc_env_playout_step(self._h, seed, s.ctypes.data, max_plies,
                                          self.n if game_stride is None else game_stride, self._which(outputs), k))

    def sync(self) -> None:
        check(self.L.fpc_env_sync(self._h))

    @property
    def stream(self) -> int:
        return self.L.fpc_env_stream(self._h) or 0

    def dlpack(self, name: str = "planes"):
        """PyCapsule("dltensor") over the library's own DLManagedTensor of one store tensor.  The consumer takes ownership
        (renames the capsule and calls the deleter); a capsule nobody consumes keeps the store alive."""
        m = self.L.fpc_env_dlpack(self._h, _NAMES[name])
        if not m:
            raise _lib.FpcError(self.L.fpc_last_error().decode())
        return _new_capsule(m, _CAPSULE_NAME, None)

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.L.fpc_env_destroy(self._h)  # the store itself goes when the last exported tensor is released
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

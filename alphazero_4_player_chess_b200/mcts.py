"""BatchedMCTS: the reference's `MCTS.search` (`src/py/mcts.py:17-43`) over `fpchess::Node`
(`src/cpp/node.{h,cpp}`) for N games at once, every tree resident in HBM.

Host-side mirror of the reference interface: `MCTS(gameType, neural_net, args).search(games)` becomes
`BatchedMCTS(R, n_games, neural_net, args).search(root_boards)`.  Per simulation the host issues
`fpc_tree_select` (descend + leaf rules + planes), the network forward (PyTorch, untouched) and
`fpc_tree_expand_backup` (softmax -> un-rotate -> mask -> renormalise -> backup -> expand); nothing
crosses to the host inside the loop, where the reference moves the dense policy to the CPU every
simulation (`mcts.py:83-87`).  PyTorch owns the device arrays; all tree logic runs in libfpc.so."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import FPC_MAX_MOVES, TreeDesc, check
from .geometry import GEOMETRIES, NUM_STATE_CHANNELS

ERR_NODE_CAP, ERR_BOARD_CAP, ERR_NO_CHILD, ERR_MOVE = 1, 2, 4, 8


class BatchedMCTS:
    def __init__(self, R: int, n_games: int, neural_net, args: dict, device: str | torch.device = "cuda",
                 batch_rotation: bool = False, node_cap: int | None = None, cuda_graph: bool = False):
        """args: the reference's dict (`alphazero.py:291-306`): uses "C" and "num_searches".
        batch_rotation=True reproduces the reference bit for bit (a whole leaf batch is rotated by the
        colour of its first expandable state); False rotates every leaf by its own side to move.
        cuda_graph=True captures one whole simulation (select -> network -> expand/backup) in a CUDA graph after
        three eager ones and replays it: the ~70 launches of a simulation become one."""
        self.geom = GEOMETRIES[R]
        self.R, self.n = R, int(n_games)
        self.net = neural_net
        self.args = args
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.FpcError("BatchedMCTS needs a CUDA device (there is no CPU fallback)")
        self.batch_rotation = bool(batch_rotation)
        self.cuda_graph = bool(cuda_graph)
        self._graph = None
        self.L = _lib.lib()
        sims = int(args["num_searches"])
        # a leaf adds at most FPC_MAX_MOVES children; the legal-move average is ~19-38 at 14x14
        self.node_cap = int(node_cap) if node_cap else 1 + sims * 96
        self.board_cap = sims + 2
        n, nc, dev = self.n, self.node_cap, self.device
        i32 = dict(dtype=torch.int32, device=dev)
        self.parent = torch.empty((n, nc), **i32)
        self.first_child = torch.empty((n, nc), **i32)
        self.n_children = torch.empty((n, nc), **i32)
        self.visits = torch.empty((n, nc), **i32)
        self.move_flat = torch.empty((n, nc), **i32)
        self.board_idx = torch.empty((n, nc), **i32)
        self.value_sum = torch.empty((n, nc), dtype=torch.float64, device=dev)
        self.prior = torch.empty((n, nc), dtype=torch.float32, device=dev)
        self.n_nodes = torch.zeros(n, **i32)
        self.n_boards = torch.zeros(n, **i32)
        self.leaf = torch.full((n,), -1, **i32)
        self.dropped = torch.zeros(n, **i32)
        self.error = torch.zeros(n, **i32)
        rec = self.geom.record_bytes
        self.boards = torch.empty((n, self.board_cap, rec), dtype=torch.uint8, device=dev)
        self.leaf_boards = torch.zeros((n, rec), dtype=torch.uint8, device=dev)
        self.leaf_flat = torch.zeros((n, FPC_MAX_MOVES), **i32)
        self.leaf_counts = torch.zeros(n, **i32)
        self.leaf_status = torch.zeros(n, **i32)
        self.k = torch.zeros(n, **i32)
        self.planes = torch.empty((n, NUM_STATE_CHANNELS, R, R), dtype=torch.float32, device=dev)
        with torch.cuda.device(self.device):
            self._track = self.L.fpc_dense_track_create(R, n)  # lets select() update self.planes in place
        if not self._track:
            raise _lib.FpcError(self.L.fpc_last_error().decode())
        d = TreeDesc()
        d.R, d.n_games, d.node_cap, d.board_cap, d.C = R, n, self.node_cap, self.board_cap, float(args["C"])
        for name in ("parent", "first_child", "n_children", "visits", "move_flat", "board_idx", "value_sum", "prior",
                     "n_nodes", "n_boards", "leaf", "dropped", "error", "boards", "leaf_boards", "leaf_flat",
                     "leaf_counts", "leaf_status", "k"):
            setattr(d, name, getattr(self, name).data_ptr())
        self.desc = d

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # ---- the three device steps ----------------------------------------------------------------
    def reset(self, root_boards: torch.Tensor) -> None:
        roots = root_boards.to(self.device, torch.uint8).contiguous()
        assert roots.shape == (self.n, self.geom.record_bytes)
        with torch.cuda.device(self.device):
            check(self.L.fpc_tree_reset(C.byref(self.desc), roots.data_ptr(), self._stream()))
        self._roots = roots  # keep alive until the kernel has run

    def select(self) -> torch.Tensor:
        """Node.ChooseLeaf for every live game; returns the encoded leaf batch [n,24,R,R]."""
        # the planes tensor is resident and only read by the network between selects: after the first full encode
        # through this instance's handle it is updated in place (FPC_FLAG_INCREMENTAL)
        with torch.cuda.device(self.device):
            check(self.L.fpc_tree_select(C.byref(self.desc), int(self.batch_rotation), self.planes.data_ptr(),
                                         _lib.FLAG_INCREMENTAL, self._track, self._stream()))
        return self.planes

    def close(self) -> None:
        if getattr(self, "_track", None):
            self.L.fpc_dense_track_destroy(self._track)
            self._track = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def expand_backup(self, logits: torch.Tensor, values: torch.Tensor) -> None:
        logits = logits.to(torch.float32).contiguous()
        values = values.to(torch.float32).reshape(-1).contiguous()
        assert logits.shape == (self.n, self.geom.action_space_size) and values.shape == (self.n,)
        with torch.cuda.device(self.device):
            check(self.L.fpc_tree_expand_backup(C.byref(self.desc), logits.data_ptr(), values.data_ptr(),
                                                self._stream()))

    # ---- MCTS.search (mcts.py:17-43) -------------------------------------------------------------
    def simulate(self) -> None:
        """One simulation for every live game (`MCTS.step` + `MCTS.expand`, mcts.py:59-89)."""
        planes = self.select()
        logits, value = self.net(planes)
        self.expand_backup(logits, value)

    @torch.no_grad()
    def search(self, root_boards: torch.Tensor, record=None, check: bool = True) -> "BatchedMCTS":
        """record: optional callable(sim, logits, value) called after every network forward (eager mode only), e.g.
        to keep the outputs of some games for a replay on the CPU oracle.  check: raise on tree errors (arena full ...)
        instead of leaving them in `self.error`."""
        self.reset(root_boards)
        sims = int(self.args["num_searches"])
        done = 0
        if self.cuda_graph and record is None:
            # eager warm-up (the first select writes the planes in full; cuDNN / cuBLAS pick their kernels), then capture
            while self._graph is None and done < min(3, sims):
                self.simulate()
                done += 1
            if self._graph is None and done < sims:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.simulate()
                self._graph = g  # capturing runs nothing: the simulation count is unchanged
            while done < sims:
                self._graph.replay()
                done += 1
        else:
            for sim in range(sims):
                planes = self.select()
                logits, value = self.net(planes)
                if record is not None:
                    record(sim, logits, value)
                self.expand_backup(logits, value)
        if check:
            self.check_errors()
        return self

    # ---- results -----------------------------------------------------------------------------------
    def check_errors(self) -> None:
        e = int(self.error.max().item())
        if e:
            raise _lib.FpcError(f"tree error bits {e:#x} (1 node_cap, 2 board_cap, 4 no child selectable, 8 move)")

    def root_children(self):
        """(flat [n,FPC_MAX_MOVES] i32, visits [n,FPC_MAX_MOVES] i32, prior f32, count [n]) of the root's
        children, zero-padded: what `AlphaZero.play` reads through GetChildren / GetVisitCount /
        GetMoveMade (`alphazero.py:104-110`)."""
        cnt = self.n_children[:, 0]
        fc = self.first_child[:, 0].long()
        idx = fc[:, None] + torch.arange(FPC_MAX_MOVES, device=self.device)[None, :]
        valid = torch.arange(FPC_MAX_MOVES, device=self.device)[None, :] < cnt[:, None]
        idx = torch.where(valid, idx, torch.zeros_like(idx))
        zero = torch.zeros((), dtype=torch.int32, device=self.device)
        flat = torch.where(valid, torch.gather(self.move_flat, 1, idx), zero)
        visits = torch.where(valid, torch.gather(self.visits, 1, idx), zero)
        prior = torch.where(valid, torch.gather(self.prior, 1, idx), torch.zeros((), device=self.device))
        return flat, visits, prior, cnt

    def action_probs(self) -> torch.Tensor:
        """`alphazero.py:104-110`: visit counts of the root's children scattered over the action space and
        normalised, [n, A*R*R] f32."""
        flat, visits, _, _ = self.root_children()
        probs = torch.zeros((self.n, self.geom.action_space_size), dtype=torch.float32, device=self.device)
        probs.scatter_add_(1, flat.long(), visits.float())
        return probs / probs.sum(dim=1, keepdim=True).clamp_min(1e-30)

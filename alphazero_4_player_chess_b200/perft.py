"""perft on the device (BASELINE.json configs[0]: perft from the standard start position).

Level-synchronous frontier expansion: one `fpc_observe` launch lists the legal moves of every frontier
board, one `fpc_make_moves` launch makes all children.  The reference counts the same tree by
copy-make recursion over `GetLegalMoves` + `TakeAction` (its `UndoMove` is one level deep,
`/root/reference/src/cpp/engine/board.h:703`).  Known answers: SURVEY.md 8c."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import FPC_MAX_MOVES, check
from .geometry import GEOMETRIES


def perft(R: int, root_record, depth: int, device: str | torch.device = "cuda", chunk: int = 1 << 20) -> list[int]:
    """Leaf counts of depths 1..depth from `root_record` ([record] bytes)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.FpcError("perft needs a CUDA device (there is no CPU fallback)")
    L = _lib.lib()
    rec = GEOMETRIES[R].record_bytes
    frontier = torch.as_tensor(np.ascontiguousarray(root_record), dtype=torch.uint8).reshape(1, rec).to(dev)
    out = []
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        for d in range(depth):
            total, nxt = 0, []
            for lo in range(0, frontier.shape[0], chunk):  # bounded workspace: 8 B x 300 moves per board
                part = frontier[lo: lo + chunk]
                n = part.shape[0]
                counts = torch.zeros(n, dtype=torch.int32, device=dev)
                last = d == depth - 1
                moves = None if last else torch.empty((n, FPC_MAX_MOVES), dtype=torch.int64, device=dev)
                check(L.fpc_observe(R, part.data_ptr(), n, None if last else moves.data_ptr(), None, counts.data_ptr(),
                                    None, None, None, -1, None, 0, stream))
                total += int(counts.sum().item())
                if last:
                    continue
                c = counts.long()
                idx = torch.repeat_interleave(torch.arange(n, device=dev), c)
                within = torch.arange(idx.numel(), device=dev) - (torch.cumsum(c, 0) - c)[idx]
                mv = moves[idx, within].contiguous()
                parents = part[idx].contiguous()
                children = torch.empty_like(parents)
                err = torch.zeros(idx.numel(), dtype=torch.int32, device=dev)
                check(L.fpc_make_moves(R, parents.data_ptr(), mv.data_ptr(), idx.numel(), children.data_ptr(),
                                       err.data_ptr(), stream))
                if int(err.abs().sum().item()):
                    raise _lib.FpcError("perft: make-move failed on a generated move")
                nxt.append(children)
            out.append(total)
            if d < depth - 1:
                frontier = torch.cat(nxt) if nxt else frontier[:0]
    return out

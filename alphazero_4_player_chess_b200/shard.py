"""Sharding of independent games over the GPUs of one box.

Games never interact, so there is no data-path collective: rank r owns the slots of its own board
store and gives them the global game ids r*n + slot, advancing by world*n whenever a slot is
re-seeded.  The playout random stream is a function of (seed, global game id, ply), which makes every
game identical whatever the number of GPUs.  The only collective is a sum of a few counters for
reporting (NCCL on GPUs; any torch.distributed backend works)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    games_per_gpu: int

    @property
    def first_game(self) -> int:
        return self.rank * self.games_per_gpu

    @property
    def game_stride(self) -> int:
        return self.world * self.games_per_gpu

    def game_id(self, slot: int, generation: int) -> int:
        """Global id of the generation-th game played in a slot of this rank."""
        return self.first_game + slot + generation * self.game_stride

    def owner(self, game_id: int) -> tuple[int, int, int]:
        """(rank, slot, generation) of a global game id."""
        generation, rest = divmod(game_id, self.game_stride)
        rank, slot = divmod(rest, self.games_per_gpu)
        return rank, slot, generation


def reduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """Sum the per-rank statistics (positions, finished games, result histogram, ...) over all ranks."""
    out = counters.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

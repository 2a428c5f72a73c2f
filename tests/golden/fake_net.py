"""TEST INFRASTRUCTURE.  A deterministic stand-in for the reference's ResNet (`src/py/net.py`) with
the same call contract (`encoded [B,24,R,R] f32 -> (logits [B, A*R*R] f32, value [B,1] f32)`, a
`.device` attribute), used to drive the reference's `MCTS.search` when the golden fixtures are made
and our PUCT path in the tests.  Every output is an exact function of the 0/1 input planes computed
in integer arithmetic and scaled by powers of two, so it is bit-identical on any host or device."""
from __future__ import annotations

import torch


class FakeNet:
    def __init__(self, R: int, device: str = "cpu"):
        self.R = R
        self.S = 24 * R * R
        self.ASZ = (8 * R + 8) * R * R
        self.device = torch.device(device)
        self.w = ((torch.arange(self.S, dtype=torch.int64, device=self.device) * 40503 + 12345) % 65521)
        self.j = 2 * torch.arange(self.ASZ, dtype=torch.int64, device=self.device) + 1
        self.calls = 0
        self.positions = 0

    def signature(self, encoded: torch.Tensor) -> torch.Tensor:
        b = encoded.shape[0]
        return (encoded.reshape(b, -1).to(torch.int64) * self.w).sum(dim=1) % (1 << 31)

    def __call__(self, encoded: torch.Tensor):
        self.calls += 1
        self.positions += int(encoded.shape[0])
        sig = self.signature(encoded)
        t = ((sig[:, None] + 1) * self.j[None, :]) % (1 << 31)
        logits = (((t * 40503) >> 7) & 0xFFFF).to(torch.float32) / 8192.0 - 4.0
        value = ((sig * 48271) % 65536).to(torch.float32) / 32768.0 - 1.0
        return logits, value[:, None]

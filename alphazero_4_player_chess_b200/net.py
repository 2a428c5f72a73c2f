"""The consumer of the encoded planes: a PyTorch policy/value network with the architecture of the
reference's `ResNet` (`src/py/net.py:6-63`: 3x3 conv stem, N residual blocks, policy head conv -> BN ->
ReLU -> Linear(A*R*R -> A*R*R), value head conv -> BN -> ReLU -> Linear -> tanh).  The network stays in
PyTorch by design (BASELINE.json north star); this module exists so that the self-play benchmark has
a random-init model of the named shape to drive.  It is not part of the hot path being replaced."""
from __future__ import annotations

import torch
from torch import nn

from .geometry import GEOMETRIES, NUM_STATE_CHANNELS


def _conv_bn_relu(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU())


class _Residual(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.a = nn.Sequential(nn.Conv2d(ch, ch, 3, padding=1), nn.BatchNorm2d(ch), nn.ReLU(),
                               nn.Conv2d(ch, ch, 3, padding=1), nn.BatchNorm2d(ch))

    def forward(self, x):
        return torch.relu(self.a(x) + x)


class PolicyValueNet(nn.Module):
    def __init__(self, R: int, blocks: int = 10, hidden: int = 128, device: str | torch.device = "cuda"):
        super().__init__()
        g = GEOMETRIES[R]
        self.device = torch.device(device)
        self.stem = _conv_bn_relu(NUM_STATE_CHANNELS, hidden)
        self.tower = nn.Sequential(*[_Residual(hidden) for _ in range(blocks)])
        self.policy = nn.Sequential(_conv_bn_relu(hidden, g.num_action_channels), nn.Flatten(),
                                    nn.Linear(g.action_space_size, g.action_space_size))
        self.value = nn.Sequential(_conv_bn_relu(hidden, NUM_STATE_CHANNELS), nn.Flatten(),
                                   nn.Linear(g.state_space_size, 1), nn.Tanh())
        self.to(self.device)

    def forward(self, x):
        x = self.tower(self.stem(x))
        return self.policy(x), self.value(x)


class InferenceNet:
    """Eval-mode wrapper used inside the search loop: bf16 weights + channels_last,
    f32 logits / values out, same call contract as the reference's `self.neural_net(encoded)`."""

    def __init__(self, net: PolicyValueNet, bf16: bool = True):
        self.dtype = torch.bfloat16 if bf16 else torch.float32
        # weights converted once (autocast would re-cast the 553 M-parameter policy Linear every call)
        self.net = net.eval().to(dtype=self.dtype, memory_format=torch.channels_last)
        self.device = net.device

    @torch.no_grad()
    def __call__(self, planes: torch.Tensor):
        x = planes.to(self.dtype).contiguous(memory_format=torch.channels_last)
        logits, value = self.net(x)
        return logits.float(), value.float()


class AutocastNet:
    """Inference view of a model that is also being trained: eval-mode forward under bf16 autocast, weights
    untouched (fp32), f32 logits / values out."""

    def __init__(self, net: PolicyValueNet):
        self.net = net
        self.device = net.device

    @torch.no_grad()
    def __call__(self, planes: torch.Tensor):
        was_training = self.net.training
        self.net.eval()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, value = self.net(planes)
        self.net.train(was_training)
        return logits.float(), value.float()

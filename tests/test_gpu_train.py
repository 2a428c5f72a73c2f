"""-m gpu: one iteration of the outer loop (train.py: self-play -> replay -> optimiser steps -> validation)
with a tiny network at 8x8: the loop runs end to end on the device and the optimiser changes the weights."""
import pytest
import torch

from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.net import AutocastNet, PolicyValueNet
from alphazero_4_player_chess_b200.selfplay import SelfPlay
from alphazero_4_player_chess_b200.train import Learner

pytestmark = pytest.mark.gpu


def test_one_learning_iteration():
    torch.manual_seed(0)
    R = 8
    model = PolicyValueNet(R, blocks=1, hidden=16, device="cuda")
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4)
    args = {"C": 3, "num_searches": 8, "temperature": 1.1, "max_game_length": 10, "heuristic_weight": 0.02,
            "batch_size": 32, "replay_buffer_capacity": 4096, "validation_buffer_capacity": 1024,
            "num_iterations": 1, "num_games": 32, "num_parallel_games": 32}
    sp = SelfPlay(R, 32, AutocastNet(model), args, start_record("EIGHT_SIMPLE"))
    learner = Learner(sp, model, opt, args)
    before = [p.detach().clone() for p in model.parameters()]
    out = learner.learn(1)
    assert out[0]["train_steps"] >= 1 and out[0]["replay"] >= 32
    assert all(torch.isfinite(torch.tensor([e["policy_loss"], e["value_loss"]])).all() for e in learner.log)
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))
    assert out[0]["validation"] is None or out[0]["validation"]["validation_loss"] > 0
    boards, flat, visits, value = learner.train_buf.sample(8)
    assert boards.shape == (8, 80) and bool((visits.sum(dim=1) > 0).all()) and bool((value.abs() <= 1.0).all())

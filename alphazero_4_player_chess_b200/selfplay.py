"""SelfPlay: the reference's self-play loop `AlphaZero.play` (`src/py/alphazero.py:81-178`) with every
game, tree and replay entry resident on the device.

Per move, for all games at once: `BatchedMCTS.search` -> visit counts of the root's children
(`alphazero.py:104-110`) -> temperature + multinomial sampling (`:114-118`) -> index-built make-move
(`state.TakeAction(Move(action_index))`, `:119-121`) -> `GetGameResult` (`:123`) -> terminal rewards by
team (`:128-137`, `handle_terminal_state :53-79`); games still running at `max_game_length` are scored
by `CalculateHeuristic * heuristic_weight` (`:161-175`).  Replay entries are device tensors (board
record + sparse visit distribution + value) instead of `MemoryEntry` copies of whole boards; the
training-side encoding `GetEncodedState(entry.state)` is produced on demand by `encoded_states()`.

This is the caller of the hot path, not part of the bit-exact contract: sampling uses
torch's device RNG.  The deterministic pieces (advance, rewards) are checked against the oracle in
tests/test_gpu_selfplay.py."""
from __future__ import annotations

import torch

from ._lib import FPC_MAX_MOVES
from .env import BatchedEnv
from .geometry import GEOMETRIES
from .mcts import BatchedMCTS


class SelfPlay:
    def __init__(self, R: int, n_games: int, neural_net, args: dict, start_record, device="cuda",
                 batch_rotation: bool = False):
        """args: the reference's dict (`alphazero.py:291-306`); uses C, num_searches, temperature,
        max_game_length, heuristic_weight."""
        self.geom = GEOMETRIES[R]
        self.R, self.n, self.args = R, int(n_games), args
        self.device = torch.device(device)
        self.env = BatchedEnv(R, self.n, device=self.device)
        self.mcts = BatchedMCTS(R, self.n, neural_net, args, device=self.device, batch_rotation=batch_rotation)
        self.start = torch.as_tensor(start_record, dtype=torch.uint8)
        T = int(args["max_game_length"])
        i32 = dict(dtype=torch.int32, device=self.device)
        self.hist_boards = torch.zeros((T, self.n, self.geom.record_bytes), dtype=torch.uint8, device=self.device)
        self.hist_flat = torch.zeros((T, self.n, FPC_MAX_MOVES), **i32)
        self.hist_visits = torch.zeros((T, self.n, FPC_MAX_MOVES), **i32)
        self.hist_valid = torch.zeros((T, self.n), dtype=torch.bool, device=self.device)
        self.hist_action = torch.full((T, self.n), -1, **i32)

    # ---- one move for every running game -----------------------------------------------------------
    def sample_actions(self, flat: torch.Tensor, visits: torch.Tensor, running: torch.Tensor) -> torch.Tensor:
        """`alphazero.py:108-118`: probs = visits / sum; probs ** (1/temperature), renormalised; multinomial."""
        w = visits.float()
        w = w / w.sum(dim=1, keepdim=True).clamp_min(1e-30)
        w = torch.pow(w, 1.0 / float(self.args["temperature"]))
        has = (w.sum(dim=1) > 0) & running
        w = torch.where(has[:, None], w, torch.ones_like(w))  # multinomial needs a non-zero row
        pick = torch.multinomial(w, 1).squeeze(1)
        action = torch.gather(flat, 1, pick[:, None]).squeeze(1)
        return torch.where(has, action, torch.full_like(action, -1))

    def advance(self, actions: torch.Tensor) -> torch.Tensor:
        """TakeAction(Move(action_index)) for every game with action >= 0 (the others keep their board), then
        GetGameResult of the new positions.  Returns the result codes [n] (0 = IN_PROGRESS)."""
        self.env.make_index(actions.to(torch.int32).contiguous())  # action -1: FPC_ERR_MOVE, record untouched
        self.env.observe(planes=False, mask=False)
        return self.env.status & 3

    @torch.no_grad()
    def play(self) -> dict:
        n, T = self.n, int(self.args["max_game_length"])
        off_turn = self.geom.off_turn
        self.env.load(self.start.numpy())
        running = torch.ones(n, dtype=torch.bool, device=self.device)
        losing_team = torch.full((n,), -1, dtype=torch.int64, device=self.device)
        self.hist_valid.zero_()
        for t in range(T):
            self.mcts.reset(self.env.boards)
            self.mcts.dropped.copy_((~running).int())  # finished games take no part in the search
            for _ in range(int(self.args["num_searches"])):
                planes = self.mcts.select()
                logits, value = self.mcts.net(planes)
                self.mcts.expand_backup(logits, value)
            flat, visits, _, _ = self.mcts.root_children()
            self.hist_boards[t].copy_(self.env.boards)
            self.hist_flat[t].copy_(flat)
            self.hist_visits[t].copy_(visits)
            actions = self.sample_actions(flat, visits, running)
            moved = running & (actions >= 0)
            self.hist_valid[t].copy_(moved)
            self.hist_action[t].copy_(actions)
            mover_team = (self.env.boards[:, off_turn] & 1).long()
            result = self.advance(actions)
            ended = moved & (result != 0)
            # alphazero.py:128: losing_team = state.GetTurn().GetTeam() of the state the move was made FROM
            losing_team = torch.where(ended, mover_team, losing_team)
            running = running & ~ended & moved
            if not bool(running.any()):
                break
        self.mcts.check_errors()
        return self.finish(running, losing_team)

    # ---- rewards (handle_terminal_state, alphazero.py:53-79,128-137,161-175) ------------------------
    def finish(self, running: torch.Tensor, losing_team: torch.Tensor) -> dict:
        off_turn = self.geom.off_turn
        entry_team = (self.hist_boards[:, :, off_turn] & 1).long()  # [T, n] team of the side to move
        # finished games: +1 for the team that is not the losing team, GetOpponentValue(1) = -1 otherwise
        win = torch.where(entry_team != losing_team[None, :], 1.0, -1.0)
        # games cut off at max_game_length: heuristic of the team to move in the final position
        curr_team = (self.env.boards[:, off_turn] & 1).long()
        heur = self.env.heuristic().float() * float(self.args["heuristic_weight"])
        cut = torch.where(entry_team == curr_team[None, :], heur[None, :], -heur[None, :])
        value = torch.where(running[None, :], cut, win)
        sel = self.hist_valid
        return {
            "boards": self.hist_boards[sel],        # [M, record] the replay states
            "child_flat": self.hist_flat[sel],      # [M, FPC_MAX_MOVES] flat action indices of the root's children
            "child_visits": self.hist_visits[sel],  # [M, FPC_MAX_MOVES] their visit counts (policy target = normalised)
            "value": value[sel],                    # [M]
            "finished": ~running,
            "losing_team": losing_team,
        }

    def policy_targets(self, replay: dict) -> torch.Tensor:
        """Dense [M, A*R*R] action-probability targets (`alphazero.py:104-110`)."""
        m = replay["child_flat"].shape[0]
        probs = torch.zeros((m, self.geom.action_space_size), dtype=torch.float32, device=self.device)
        probs.scatter_add_(1, replay["child_flat"].long(), replay["child_visits"].float())
        return probs / probs.sum(dim=1, keepdim=True).clamp_min(1e-30)

    def encoded_states(self, replay: dict) -> torch.Tensor:
        """`GetEncodedState(entry.state)` for every replay entry (`alphazero.py:71-75`): each state rotated by its
        own side to move, [M,24,R,R] f32."""
        boards = replay["boards"].contiguous()
        env = BatchedEnv(self.R, boards.shape[0], device=self.device)
        env.load(boards)
        return env.encode(k=-1)

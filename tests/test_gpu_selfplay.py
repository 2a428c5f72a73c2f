"""-m gpu: the device-resident self-play driver (selfplay.py, the reference's AlphaZero.play).  Sampling
uses torch's RNG, so the check is a replay: the oracle re-plays every game from the recorded actions
and must see the same boards, results and rewards."""
import numpy as np
import pytest
import torch

from alphazero_4_player_chess_b200.fen import start_record
from alphazero_4_player_chess_b200.geometry import GEOMETRIES
from alphazero_4_player_chess_b200.selfplay import SelfPlay
from tests.golden.fake_net import FakeNet
from tests.util import oracle_for

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,R,n,T", [("EIGHT_SIMPLE", 8, 24, 40), ("STANDARD", 14, 8, 6)])
def test_selfplay_replays_on_the_oracle(name, R, n, T):
    torch.manual_seed(1234)
    g = GEOMETRIES[R]
    o = oracle_for(R)
    args = {"C": 3, "num_searches": 12, "temperature": 1.1, "max_game_length": T, "heuristic_weight": 0.02}
    sp = SelfPlay(R, n, FakeNet(R, device="cuda"), args, start_record(name))
    replay = sp.play()
    torch.cuda.synchronize()
    boards, valid, action = sp.hist_boards.cpu().numpy(), sp.hist_valid.cpu().numpy(), sp.hist_action.cpu().numpy()
    flat, visits = sp.hist_flat.cpu().numpy(), sp.hist_visits.cpu().numpy()
    finished, losing = replay["finished"].cpu().numpy(), replay["losing_team"].cpu().numpy()
    want_values, n_entries = [], 0
    final = sp.env.boards.cpu().numpy()
    for game in range(n):
        rec = start_record(name)
        ended = False
        entries = []
        for t in range(T):
            if not valid[t, game]:
                break
            assert np.array_equal(boards[t, game], rec), (game, t)
            legal = sorted(set(o.move_flat_index(m) for m in o.legal_moves(rec)))
            kids = [f for f, v in zip(flat[t, game], visits[t, game]) if v > 0]
            assert set(kids) <= set(legal) and int(action[t, game]) in legal
            entries.append(int(rec[g.off_turn]) & 1)
            mover_team = int(rec[g.off_turn]) & 1
            rec = o.make_index(rec, int(action[t, game]))
            res, _, _ = o.game_result(rec)
            if res != 0:
                ended = True
                assert finished[game] and losing[game] == mover_team
                break
        assert np.array_equal(final[game], rec)
        if ended:
            vals = [1.0 if team != losing[game] else -1.0 for team in entries]
        else:
            assert not finished[game] or not entries
            cur = int(rec[g.off_turn]) & 1
            h = np.float32(o.heuristic(rec, cur)) * np.float32(0.02)
            vals = [float(h) if team == cur else float(-h) for team in entries]
        want_values.append(vals)
        n_entries += len(entries)
    assert replay["boards"].shape[0] == n_entries and n_entries > 0
    # replay entries are ordered by (ply, game)
    got = replay["value"].cpu().numpy()
    order = [(t, game) for t in range(T) for game in range(n) if valid[t, game]]
    per_game_pos = {game: 0 for game in range(n)}
    for i, (t, game) in enumerate(order):
        assert abs(got[i] - want_values[game][per_game_pos[game]]) < 1e-6
        per_game_pos[game] += 1
    pol = sp.policy_targets(replay)
    assert torch.allclose(pol.sum(dim=1), torch.ones(n_entries, device="cuda"), atol=1e-5)
    enc = sp.encoded_states(replay).cpu().numpy()
    recs = replay["boards"].cpu().numpy()
    assert np.array_equal(enc, o.encode(recs, recs[:, g.off_turn].astype(np.int32)))

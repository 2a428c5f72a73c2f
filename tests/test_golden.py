"""CPU: the restatement oracle against the committed golden fixtures (tests/golden/*.npz), which
were produced by the reference itself -- the unmodified rules engine and the reference's own pybind
module driven by its own Python (generator: tests/golden/make_golden.py).  These run everywhere,
including where /root/reference and oracle/_ref do not exist."""
import os

import numpy as np
import pytest

from alphazero_4_player_chess_b200.fen import START_FENS, start_record
from alphazero_4_player_chess_b200.geometry import GEOMETRIES
from tests.util import SEED, oracle_for

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENGINE = ["STANDARD", "THIRTEEN", "TEN", "EIGHT", "EIGHT_SIMPLE"]


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def unpack(bits, shape):
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].reshape(shape).astype(np.float32)


@pytest.mark.parametrize("name", ENGINE)
@pytest.mark.parametrize("tag", ["c0", "c1"])
def test_engine_fixtures(name, tag):
    _, R = START_FENS[name]
    o = oracle_for(R)
    z = load(f"engine_{name}.npz")
    start = start_record(name, castling=tag == "c1")
    assert int(z["seed"]) == SEED
    for d, want in enumerate(z[f"perft_{tag}"], start=1):
        if want <= 200000:
            assert o.perft(start, d) == int(want)
    recs, game = z[f"recs_{tag}"], z[f"game_{tag}"]
    legal, off = z[f"legal_{tag}"], z[f"legal_off_{tag}"]
    early = 0
    for g in np.unique(game):
        sel = np.nonzero(game == g)[0]
        p = o.playout(start, SEED, int(g), int(z["max_plies"]))
        assert p["n"] == len(sel)
        assert np.array_equal(p["recs"], recs[sel])
        assert np.array_equal(p["n_legal"], z[f"n_legal_{tag}"][sel])
        assert np.array_equal(p["result"], z[f"result_{tag}"][sel])
        assert np.array_equal(p["moves"], z[f"moves_{tag}"][sel])
    for i, rec in enumerate(recs):
        want = legal[off[i]: off[i + 1]]
        got = o.legal_moves(rec)
        assert np.array_equal(got, want), i
        res, nl, kc = o.game_result(rec)
        ref_res = int(z[f"result_ref_{tag}"][i])
        if ref_res != res:  # GetGameResult's order-dependent early-out
            assert res == 0 and kc and nl > 0
            early += 1
        if len(want):
            m = int(want[-1])
            assert np.array_equal(o.make_move(rec, m), z[f"after_last_legal_{tag}"][i])
            assert np.array_equal(o.make_index(rec, o.move_flat_index(m)), z[f"after_last_legal_index_{tag}"][i])
    assert early <= len(recs) // 50


@pytest.mark.parametrize("R", [14, 8])
def test_binding_fixtures(R):
    g = GEOMETRIES[R]
    o = oracle_for(R)
    z = load(f"binding_R{R}.npz")
    st = z["statics"]
    assert list(st[:4]) == [24, g.state_space_size, g.num_action_channels, g.action_space_size]
    assert list(st[4:]) == [R - 1, 8 * (R - 1), 8, g.IA]
    # FEN loader == the reference's fen_parser + Board ctor
    for key in z.files:
        if key.startswith("start_") and not key.endswith("_str"):
            assert np.array_equal(start_record(key[len("start_"):], castling=False), z[key]), key
    recs = z["recs"]
    n = len(recs)
    turns = recs[:, g.off_turn].astype(np.int32)
    # encoder: per-state rotation, and whole batches rotated by the colour of states[0]
    assert np.array_equal(o.encode(recs, turns), unpack(z["planes_own"], (n, 24, R, R)))
    for k in range(4):
        order = z[f"planes_batch_k{k}_order"]
        assert turns[order[0]] == k
        assert np.array_equal(o.encode(recs[order], k), unpack(z[f"planes_batch_k{k}"], (len(order), 24, R, R)))
    # legal mask (absolute coordinates) and legal flat indices
    assert np.array_equal(o.mask(recs), unpack(z["mask"], (n, g.num_action_channels, R, R)))
    off = z["legal_off"]
    for i, rec in enumerate(recs):
        lm = o.legal_moves(rec)
        flat = [o.move_flat_index(m) for m in lm]
        assert flat == z["legal_flat"][off[i]: off[i + 1]].tolist(), i
        assert sorted((f, (int(m) >> 8) & 0xff) for f, m in zip(flat, lm)) == sorted(
            zip(z["legal_flat"][off[i]: off[i + 1]].tolist(), z["legal_to"][off[i]: off[i + 1]].tolist()))
        res, nl, kc = o.game_result(rec)
        if res != int(z["result"][i]):
            assert res == 0 and kc and nl > 0
        assert [o.heuristic(rec, 0), o.heuristic(rec, 1)] == z["heuristic"][i].tolist()
        if len(lm):
            last = flat[-1]
            cands = [o.make_move(rec, m) for m, f in zip(lm, flat) if f == last]
            assert any(np.array_equal(c, z["after_full"][i]) for c in cands), i
            assert np.array_equal(o.make_index(rec, last), z["after_index"][i]), i
    # ParseActionspace = rot90 by -colour on the spatial dims, planes untouched
    A = g.num_action_channels
    for k in range(4):
        ar = np.arange(A * R * R, dtype=np.int32).reshape(A, R, R)
        assert np.array_equal(np.rot90(ar, -k, axes=(1, 2)).reshape(-1), z["parse_actionspace_perm"][k])


def test_known_answers_from_the_survey():
    """EIGHT_SIMPLE encoder / mask non-zeros and the 14x14 start flat indices."""
    o8, o14 = oracle_for(8), oracle_for(14)
    rec = start_record("EIGHT_SIMPLE")
    enc = o8.encode(rec[None], 0)[0]
    want = {(2, 7, 2), (2, 7, 5), (4, 7, 4), (5, 4, 1), (10, 4, 0), (11, 1, 3), (11, 1, 4), (11, 1, 5), (14, 0, 2),
            (14, 0, 5), (16, 0, 3), (17, 3, 6), (22, 3, 7), (23, 6, 2), (23, 6, 3), (23, 6, 4)}
    assert set(map(tuple, np.argwhere(enc == 1).tolist())) == want
    m = o8.mask(rec[None])[0]
    want = {(0, 6, 2), (0, 6, 3), (0, 6, 4), (0, 7, 5), (1, 6, 2), (1, 6, 3), (1, 6, 4), (1, 7, 5), (2, 7, 5), (3, 7, 5),
            (4, 7, 5), (14, 7, 4), (42, 7, 2), (49, 7, 4)}
    assert set(map(tuple, np.argwhere(m == 1).tolist())) == want
    rec = start_record("STANDARD")
    flat = sorted(o14.move_flat_index(mv) for mv in o14.legal_moves(rec))
    assert flat == list(range(171, 179)) + list(range(367, 375)) + [20962, 20967, 21354, 21359]
    enc = o14.encode(rec[None], 0)[0]
    assert enc.sum(axis=(1, 2)).astype(int).tolist() == [2, 2, 2, 1, 1, 8] * 4


@pytest.mark.parametrize("R", [8, 14])
@pytest.mark.parametrize("case", ["a", "b"])
def test_mcts_fixtures(R, case):
    """The MCTS restatement (oracle/mcts_port.py) against the reference's own MCTS.search run on the
    reference binding with the stand-in network: identical root children and visit counts."""
    from oracle.mcts_port import search
    from tests.golden.fake_net import FakeNet
    z = load(f"mcts_R{R}.npz")
    o = oracle_for(R)
    roots = z[f"{case}_roots"]
    net = FakeNet(R)
    trees = search(o, net, roots, 3, int(z[f"{case}_sims"]), batch_rotation=True)
    off = z[f"{case}_child_off"]
    for gi, t in enumerate(trees):
        ch = t.children[0]
        assert [t.move_flat[c] for c in ch] == z[f"{case}_child_flat"][off[gi]: off[gi + 1]].tolist(), gi
        assert [t.visits[c] for c in ch] == z[f"{case}_child_visits"][off[gi]: off[gi + 1]].tolist(), gi
        assert t.visits[0] == int(z[f"{case}_root_visits"][gi])
        assert len(t.parent) == int(z[f"{case}_n_nodes"][gi])
    assert net.calls == int(z[f"{case}_nn_calls"]) and net.positions == int(z[f"{case}_nn_positions"])


@pytest.mark.parametrize("R", [14, 8])
def test_viewer_fixtures(R):
    """Attacked squares per colour / team (the pygame viewer's queries, src/cpp/board.cpp:120-232) as the reference's
    own binding reported them."""
    o = oracle_for(R)
    z = load(f"viewer_R{R}.npz")
    assert z["attack_map"].any()
    for rec, want in zip(z["recs"], z["attack_map"]):
        got = o.attack_map(rec)
        assert np.array_equal(got, want)
        for sq in range(0, R * R, 11):
            for c in range(4):
                assert o.is_attacked_by_player(rec, sq, c) == bool((want[sq] >> c) & 1)

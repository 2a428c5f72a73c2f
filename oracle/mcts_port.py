"""TEST INFRASTRUCTURE.  CPU restatement of the reference's batched MCTS: `src/py/mcts.py:17-89`
(MCTS.search / step / expand) over `src/cpp/node.cpp` (Node::ChooseLeaf :19-47, SelectChild :49-78,
Expand :79-98, Backpropagate :133-142), with boards held as board records and the rules taken from the
restatement oracle (oracle/port.py).  Pinned against the reference itself by tests/test_golden.py
(fixtures tests/golden/mcts_R*.npz, produced by the reference's own MCTS.search on its own binding).

Only tests/ may import this."""
from __future__ import annotations

import math

import numpy as np
import torch

from .port import Oracle


class Tree:
    """One game's tree; node 0 is the root (visit_count 1, `mcts.py:30`)."""

    def __init__(self, o: Oracle, C: float, root_rec: np.ndarray):
        self.o, self.C = o, float(C)
        self.parent = [-1]
        self.children = [[]]
        self.visits = [1]
        self.value_sum = [0.0]
        self.prior = [0.0]
        self.move_flat = [-1]
        self.rec = [np.ascontiguousarray(root_rec).copy()]

    # node.cpp:49-78
    def select_child(self, node: int) -> int:
        best, best_ucb = -1, -math.inf
        lg = math.log(math.sqrt(float(self.visits[node])))
        for ch in self.children[node]:
            n = self.visits[ch]
            q = self.value_sum[ch] / n if n > 0 else 0.0
            ucb = q + self.C * math.sqrt(lg / (1 + n)) * self.prior[ch]
            if ucb > best_ucb:
                best, best_ucb = ch, ucb
        if best < 0:
            raise RuntimeError("Failed to select a child.")
        return best

    # node.cpp:133-142
    def backpropagate(self, node: int, value: float) -> None:
        v = float(np.float32(value))
        while node >= 0:
            self.value_sum[node] += v
            self.visits[node] += 1
            v = -v
            node = self.parent[node]

    # node.cpp:19-47
    def choose_leaf(self):
        node = 0
        while self.children[node]:
            node = self.select_child(node)
        res, _, _ = self.o.game_result(self.rec[node])
        if res != 0:
            self.backpropagate(node, 0.0 if res == 3 else -1.0)
            return None
        return node

    # node.cpp:79-98: one child per non-zero policy entry, board = copy + MakeMove(index-built move)
    def expand(self, node: int, flats, probs) -> None:
        for flat, p in zip(flats, probs):
            self.parent.append(node)
            self.children.append([])
            self.visits.append(1)  # node.h:28 default
            self.value_sum.append(0.0)
            self.prior.append(float(p))
            self.move_flat.append(int(flat))
            self.rec.append(self.o.make_index(self.rec[node], int(flat)))
            self.children[node].append(len(self.parent) - 1)


def search(o: Oracle, net, root_recs, C: float, num_searches: int, batch_rotation: bool = True, replay=None,
           prior_source=None, prior_rtol: float = 2e-6):
    """mcts.py:17-43.  batch_rotation=True reproduces the reference (the whole leaf batch is encoded and
    un-rotated by the colour of states[0], `board.cpp:354-355`, `mcts.py:69`); False rotates every leaf
    by its own side to move (the native mode of the CUDA path).

    Record / replay: replay(sim, game_indices) -> (logits [k, A*R*R], value [k, 1]) stands in for
    the network with outputs recorded from another run, so that float noise of the network cannot leak into the
    comparison.  prior_source [n_games][node_cap] f32 does the same for the f32 priors: the priors computed here
    (torch softmax, op by op) must agree with the recorded ones within prior_rtol, and the recorded bit patterns are
    the ones stored in the tree, so that everything downstream (selection, visit counts, value sums) compares exactly."""
    R, A = o.R, o.A
    off_turn = R * R
    trees = [Tree(o, C, r) for r in root_recs]
    live = list(range(len(trees)))
    for _ in range(num_searches):
        leaves = []
        for gi in live[:]:
            leaf = trees[gi].choose_leaf()
            if leaf is None:
                live.remove(gi)  # mcts.py:22-23: the root drops out of all remaining simulations
            else:
                leaves.append((gi, leaf))
        if not leaves:
            continue
        recs = np.stack([trees[gi].rec[leaf] for gi, leaf in leaves])
        turns = recs[:, off_turn].astype(np.int32)
        k = np.full(len(leaves), turns[0], dtype=np.int32) if batch_rotation else turns
        if replay is not None:
            logits, value = replay(_, [gi for gi, _leaf in leaves])
        else:
            logits, value = net(torch.from_numpy(o.encode(recs, k)))
        flat_policy = torch.softmax(logits, dim=1)  # mcts.py:67
        pol = flat_policy.view(-1, A, R, R)
        if batch_rotation:
            pol = torch.rot90(pol, -int(turns[0]), (-2, -1))  # board.cpp:257-263
        else:
            pol = torch.stack([torch.rot90(pol[i], -int(turns[i]), (-2, -1)) for i in range(len(leaves))])
        pol = pol * torch.from_numpy(o.mask(recs))  # mcts.py:74
        pol = pol / torch.sum(pol, dim=(1, 2, 3), keepdim=True)  # mcts.py:75-76
        vals = value.squeeze(1)
        for i, (gi, leaf) in enumerate(leaves):  # node.cpp:144-154
            trees[gi].backpropagate(leaf, float(vals[i]))
        flat = pol.reshape(len(leaves), -1)
        for i, (gi, leaf) in enumerate(leaves):  # mcts.py:82-89: nonzero() order = ascending flat index
            nz = torch.nonzero(flat[i]).view(-1)
            probs = flat[i][nz].tolist()
            if prior_source is not None:
                first = len(trees[gi].parent)
                rec = prior_source[gi][first: first + len(probs)]
                np.testing.assert_allclose(rec, np.array(probs, dtype=np.float32), rtol=prior_rtol, atol=0)
                probs = [float(x) for x in rec]
            trees[gi].expand(leaf, nz.tolist(), probs)
    return trees

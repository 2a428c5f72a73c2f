// TEST INFRASTRUCTURE.  Compiles the product's device rules header (csrc/fpc_device.cuh) for the
// HOST with g++ and walks one game sequentially the way one warp does in observe_kernel, so the
// rules can be checked against the oracle in the GPU-less build container.  The warp
// choreography itself (ballots, scans, streaming stores) is only exercised by the -m gpu tests.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "fpc_device.cuh"

using namespace fpc;

template <class G>
static void load(WarpScratch<G> &s, const uint8_t *rec) {
  memset(s.mb, WALL, sizeof s.mb);
  for (int i = 0; i < 4; ++i) s.king[i] = NO_SQ, s.rights[i] = rec[G::OFF_RIGHTS + i];
  s.turn = rec[G::OFF_TURN] & 3;
  for (int sq = 0; sq < G::NSQ; ++sq) {
    int r = sq / G::R, c = sq % G::R;
    if (!G::legal(r, c)) continue;
    uint32_t p = rec[sq];
    put_cell(s.mb, G::mb(r, c), present(p) ? p : EMPTY);
    if (present(p) && type_of(p) == KING) s.king[color_of(p)] = (uint8_t)G::mb(r, c);
  }
}

template <class G>
static void store(const WarpScratch<G> &s, uint8_t *rec) {
  memset(rec, 0, G::REC);
  for (int sq = 0; sq < G::NSQ; ++sq) {
    int r = sq / G::R, c = sq % G::R;
    rec[sq] = G::legal(r, c) ? s.mb[G::mb(r, c)] : EMPTY;
  }
  rec[G::OFF_TURN] = (uint8_t)s.turn;
  for (int i = 0; i < 4; ++i) {
    rec[G::OFF_RIGHTS + i] = s.rights[i];
    rec[G::OFF_KING + i] = s.king[i] == NO_SQ ? G::NSQ : G::sq_of_mb(s.king[i]);
  }
}

template <class G>
static int legal_compact(WarpScratch<G> &s, std::vector<uint32_t> &out, int *status) {
  const int turn = s.turn;
  std::vector<uint32_t> pseudo;
  const int king_sq = s.king[turn];
  if (king_sq != NO_SQ) {
    for (int sq = 0; sq < G::NSQ; ++sq) {
      int r = sq / G::R, c = sq % G::R;
      if (!G::legal(r, c)) continue;
      int from = G::mb(r, c);
      uint32_t p = s.mb[from];
      if (!present(p) || color_of(p) != turn) continue;
      for (int line = 0; line < 4; ++line) {
        Run lo, hi;
        int kind;
        gen_item<G>(s.mb, from, line, lo, hi, kind);
        for (int j = 0; j < lo.cnt; ++j)
          pseudo.push_back(kind == 0 ? pack_compact<G>(from, from + lo.delta * (j + 1), lo.plane0 + j, NO_PIECE, 0)
                                     : pack_compact<G>(from, from + lo.delta, lo.plane0, KNIGHT + j, 0));
        for (int j = 0; j < hi.cnt; ++j) pseudo.push_back(pack_compact<G>(from, from + hi.delta * (j + 1), hi.plane0 + j, NO_PIECE, 0));
      }
    }
    for (int side = 0; side < 2; ++side) {
      uint32_t mv = gen_castle<G>(s.mb, king_sq, turn, s.rights[turn], side);
      if (mv) pseudo.push_back(mv);
    }
  }
  out.clear();
  bool takes_king = false;
  for (uint32_t mv : pseudo)
    if (king_safe_after<G>(s.mb, s.king, turn, mv)) {
      out.push_back(mv);
      uint32_t cap = s.mb[mv & 0xff];
      if (((mv >> 8) & 3) == 0 && present(cap) && type_of(cap) == KING) takes_king = true;
    }
  std::sort(out.begin(), out.end());
  const bool ry = (turn & 1) == 0;
  int result = 0, st = 0;
  if (king_sq == NO_SQ) result = ry ? 2 : 1;
  else if (out.empty()) {
    Patch none{0x1000, 0x1000, 0x1000, 0x1000, 0, 0};
    bool chk = attacked_by_team<G, false>(s.mb, 1 - (turn & 1), king_sq, none);
    result = chk ? (ry ? 2 : 1) : 3;
    if (chk) st |= 0x100;
  }
  st |= result;
  if (takes_king) st |= 0x200;
  *status = st;
  return (int)pseudo.size();
}

template <class G>
static int step(const uint8_t *rec, uint64_t seed, uint64_t game, uint64_t ply, uint8_t *out_rec, uint64_t *moves,
                int *n_legal, int *status, uint64_t *chosen, int *n_pseudo) {
  static WarpScratch<G> s;
  load<G>(s, rec);
  std::vector<uint32_t> legal;
  int np = legal_compact<G>(s, legal, status);
  if (n_pseudo) *n_pseudo = np;
  *n_legal = (int)legal.size();
  for (size_t i = 0; i < legal.size(); ++i) moves[i] = expand_move<G>(s.mb, s.rights, legal[i]);
  *chosen = 0;
  if ((*status & 3) == 0) {
    uint32_t pick = (uint32_t)(((mix64(seed, game, ply) >> 32) * (uint64_t)legal.size()) >> 32);
    *chosen = moves[pick];
    make_compact<G>(s, legal[pick]);
  }
  store<G>(s, out_rec);
  return 0;
}

extern "C" int emul_step(int R, const uint8_t *rec, uint64_t seed, uint64_t game, uint64_t ply, uint8_t *out_rec,
                         uint64_t *moves, int *n_legal, int *status, uint64_t *chosen, int *n_pseudo) {
  switch (R) {
    case 14: return step<Geo<14, 3>>(rec, seed, game, ply, out_rec, moves, n_legal, status, chosen, n_pseudo);
    case 13: return step<Geo<13, 3>>(rec, seed, game, ply, out_rec, moves, n_legal, status, chosen, n_pseudo);
    case 10: return step<Geo<10, 2>>(rec, seed, game, ply, out_rec, moves, n_legal, status, chosen, n_pseudo);
    case 8: return step<Geo<8, 2>>(rec, seed, game, ply, out_rec, moves, n_legal, status, chosen, n_pseudo);
  }
  return -1;
}

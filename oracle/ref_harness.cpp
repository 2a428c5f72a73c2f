// TEST INFRASTRUCTURE -- not product code.
//
// C-ABI harness around the UNMODIFIED reference rules engine (chess::Board,
// /root/reference/src/cpp/engine/board.{h,cpp}) and the reference action-index
// map (fpchess::Move, /root/reference/src/cpp/move.{h,cpp}).  It is compiled by
// oracle/Makefile from the reference sources where they lie; the only edit is a
// stream substitution of the three geometry constants at engine/board.h:22-24
// (the checked-in engine is 8x8; BASELINE.json's configs are 14x14).  Output
// goes to oracle/_ref/libref_engine_R<R>.so (git-ignored).
//
// The harness only calls the reference's public API; it adds no rules logic of
// its own.  Board state crosses the ABI as the "board record" defined in
// include/fpc.h (R*R piece bytes in the reference's own Piece bit layout,
// engine/board.h:101-104, then turn, 4 castling bytes, 4 king squares, pad).
//
// Used by: tests/ (oracle pinning + bulk differential tests) and bench.py's
// cpu_baseline / --impl reference legs.  Nothing under the product package
// links or loads it.

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <thread>
#include <vector>
#include <unordered_map>

#include "move.h"            // reference: src/cpp/move.h (pulls engine/board.h)

using chess::Board;
using chess::BoardLocation;
using chess::CastlingRights;
using chess::Piece;
using chess::Player;
using chess::PlayerColor;

static_assert(sizeof(chess::Move) == 8, "reference Move is an 8-byte value");
static_assert(sizeof(fpchess::Move) == 8, "adapter Move adds no fields");

namespace {

constexpr int R = chess::rows_;
constexpr int NSQ = R * R;
constexpr int REC = ((NSQ + 12 + 15) / 16) * 16;
constexpr int OFF_TURN = NSQ;
constexpr int OFF_RIGHTS = NSQ + 1;
constexpr int OFF_KING = NSQ + 5;

struct InitOnce {
  InitOnce() { fpchess::Move::InitializeMoveIndexMap(); }   // wrapper.cpp:133
} g_init;

Board BoardFromRecord(const uint8_t *rec) {
  std::unordered_map<BoardLocation, Piece> pieces;
  for (int sq = 0; sq < NSQ; ++sq) {
    uint8_t b = rec[sq];
    if (b & 0x80) {
      pieces.emplace(BoardLocation(sq / R, sq % R),
                     Piece(static_cast<PlayerColor>((b >> 5) & 3),
                           static_cast<chess::PieceType>((b >> 2) & 7)));
    }
  }
  std::unordered_map<Player, CastlingRights> rights;
  for (int c = 0; c < 4; ++c) {
    uint8_t b = rec[OFF_RIGHTS + c];
    rights.emplace(Player(static_cast<PlayerColor>(c)),
                   CastlingRights((b >> 6) & 1, (b >> 5) & 1));
  }
  return Board(Player(static_cast<PlayerColor>(rec[OFF_TURN] & 3)), pieces, rights);
}

void RecordFromBoard(const Board &b, uint8_t *rec) {
  std::memset(rec, 0, REC);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < R; ++c) {
      Piece p = b.GetPiece(r, c);
      uint8_t bits;
      std::memcpy(&bits, &p, 1);
      rec[r * R + c] = bits;
    }
  rec[OFF_TURN] = static_cast<uint8_t>(b.GetTurn().GetColor());
  auto cr = b.GetCastlingRights();
  for (int c = 0; c < 4; ++c) {
    uint8_t bits;
    std::memcpy(&bits, &cr[c], 1);
    rec[OFF_RIGHTS + c] = bits;
  }
  for (int c = 0; c < 4; ++c) {
    BoardLocation k = b.GetKingLocation(static_cast<PlayerColor>(c));
    rec[OFF_KING + c] = k.Present() ? static_cast<uint8_t>(k.GetRow() * R + k.GetCol())
                                    : static_cast<uint8_t>(NSQ);
  }
}

inline uint64_t MoveBits(const chess::Move &m) {
  uint64_t v;
  std::memcpy(&v, &m, 8);
  return v;
}
inline fpchess::Move MoveFromBits(uint64_t v) {
  fpchess::Move m;
  std::memcpy(static_cast<void *>(&m), &v, 8);
  return m;
}

// Exactly fpchess::Board::GetLegalMoves (src/cpp/board.cpp:94-118) with
// IsKingSafeAfterMove (:59-68) inlined; that class needs libtorch, the rules
// it calls are all chess::Board's.
size_t LegalMoves(Board &b, chess::Move *out) {
  chess::Move buf[300];
  size_t n = b.GetPseudoLegalMoves2(buf, 300);
  size_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    Player me = b.GetTurn();
    b.MakeMove(buf[i]);
    bool safe = !b.IsKingInCheck(me);
    b.UndoMove();
    if (safe) out[k++] = buf[i];
  }
  return k;
}

inline int SortKey(const chess::Move &m) {
  fpchess::Move fm(m);
  return fm.GetFlatIndex() * 8 + static_cast<int>(m.GetPromotionPieceType());
}

// Canonical order = ascending (flat action index, promotion type).  One
// GetFlatIndex per move (what GetLegalMovesIndices pays, src/cpp/board.cpp:438).
void SortCanonical(chess::Move *legal, size_t n) {
  std::pair<int, chess::Move> keyed[300];
  for (size_t i = 0; i < n; ++i) keyed[i] = {SortKey(legal[i]), legal[i]};
  std::sort(keyed, keyed + n, [](const auto &a, const auto &c) { return a.first < c.first; });
  for (size_t i = 0; i < n; ++i) legal[i] = keyed[i].second;
}

inline uint64_t Mix(uint64_t seed, uint64_t game, uint64_t ply) {
  uint64_t z = seed ^ (game * 0x9E3779B97F4A7C15ULL) ^ (ply * 0xBF58476D1CE4E5B9ULL);
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline uint32_t Pick(uint64_t seed, uint64_t game, uint64_t ply, uint32_t n) {
  return static_cast<uint32_t>(((Mix(seed, game, ply) >> 32) * static_cast<uint64_t>(n)) >> 32);
}

// Order-independent result: branches (i) and (iii) of
// chess::Board::GetGameResult (engine/board.cpp:895-899, 927-938) evaluated
// through the reference's own predicates.  Differs from GetGameResult only in
// the order-dependent early-out of branch (ii) (:919-922).
int CanonicalResult(const Board &b, size_t n_legal) {
  Player me = b.GetTurn();
  bool ry = me.GetTeam() == chess::RED_YELLOW;
  if (!b.GetKingLocation(me.GetColor()).Present()) return ry ? chess::WIN_BG : chess::WIN_RY;
  if (n_legal > 0) return chess::IN_PROGRESS;
  if (!b.IsKingInCheck(me)) return chess::STALEMATE;
  return ry ? chess::WIN_BG : chess::WIN_RY;
}

uint64_t Perft(const Board &b, int depth) {
  Board w(b);
  chess::Move legal[300];
  size_t n = LegalMoves(w, legal);
  if (depth <= 1) return n;
  uint64_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    Board c(w);                 // copy-make: UndoMove is one level deep (engine/board.h:703)
    c.MakeMove(legal[i]);
    total += Perft(c, depth - 1);
  }
  return total;
}

}  // namespace

extern "C" {

int ref_rows() { return R; }
int ref_invalid_area() { return chess::invalid_area; }
int ref_record_bytes() { return REC; }

int ref_pseudo_moves(const uint8_t *rec, uint64_t *out, int cap) {
  Board b = BoardFromRecord(rec);
  chess::Move buf[300];
  size_t n = b.GetPseudoLegalMoves2(buf, 300);
  for (size_t i = 0; i < n && static_cast<int>(i) < cap; ++i) out[i] = MoveBits(buf[i]);
  return static_cast<int>(n);
}

// rec_after (may be null) receives the board as GetLegalMoves leaves it.
int ref_legal_moves(const uint8_t *rec, uint64_t *out, int cap, uint8_t *rec_after) {
  Board b = BoardFromRecord(rec);
  chess::Move buf[300];
  size_t n = LegalMoves(b, buf);
  for (size_t i = 0; i < n && static_cast<int>(i) < cap; ++i) out[i] = MoveBits(buf[i]);
  if (rec_after) RecordFromBoard(b, rec_after);
  return static_cast<int>(n);
}

int ref_game_result(const uint8_t *rec) {
  Board b = BoardFromRecord(rec);
  return static_cast<int>(b.GetGameResult());
}

int ref_king_in_check(const uint8_t *rec, int color) {
  Board b = BoardFromRecord(rec);
  return b.IsKingInCheck(Player(static_cast<PlayerColor>(color))) ? 1 : 0;
}

int ref_is_attacked_by_team(const uint8_t *rec, int team, int sq) {
  Board b = BoardFromRecord(rec);
  return b.IsAttackedByTeam(static_cast<chess::Team>(team), BoardLocation(sq / R, sq % R)) ? 1 : 0;
}

int ref_heuristic(const uint8_t *rec, int team) {
  Board b = BoardFromRecord(rec);
  return b.CalculateHeuristic(static_cast<chess::Team>(team));
}

int ref_make_move(const uint8_t *rec, uint64_t move, uint8_t *out) {
  Board b = BoardFromRecord(rec);
  try {
    b.MakeMove(MoveFromBits(move));
  } catch (const std::exception &) {
    return -1;
  }
  RecordFromBoard(b, out);
  return 0;
}

// Index-built move, the self-play path: fpchess::Move(int flat_index) (src/cpp/move.cpp:41-61)
int ref_make_index(const uint8_t *rec, int flat_index, uint8_t *out) {
  Board b = BoardFromRecord(rec);
  try {
    fpchess::Move m(flat_index);
    b.MakeMove(m);
  } catch (const std::exception &) {
    return -1;
  }
  RecordFromBoard(b, out);
  return 0;
}

uint64_t ref_move_from_flat(int flat_index) { return MoveBits(fpchess::Move(flat_index)); }

int ref_move_flat_index(uint64_t move) {
  try {
    return MoveFromBits(move).GetFlatIndex();
  } catch (const std::exception &) {
    return -1;
  }
}

uint64_t ref_perft(const uint8_t *rec, int depth) {
  Board b = BoardFromRecord(rec);
  return Perft(b, depth);
}

uint64_t ref_mix(uint64_t seed, uint64_t game, uint64_t ply) { return Mix(seed, game, ply); }

// Deterministic random playout.  Per ply p the record of
// the position BEFORE the move goes to recs[p]; n_legal[p], result[p] describe
// that position (result = canonical, result_ref = GetGameResult verbatim);
// moves[p] is the move played (0 if the game ended there).
// Returns the number of positions written (<= max_plies).
int ref_playout(const uint8_t *start, uint64_t seed, uint64_t game, int max_plies,
                uint8_t *recs, int *n_legal, int *result, int *result_ref, uint64_t *moves) {
  Board b = BoardFromRecord(start);
  chess::Move legal[300];
  int p = 0;
  for (; p < max_plies; ++p) {
    if (recs) RecordFromBoard(b, recs + static_cast<size_t>(p) * REC);
    int res_ref = static_cast<int>(b.GetGameResult());
    size_t n = LegalMoves(b, legal);
    int res = CanonicalResult(b, n);
    if (n_legal) n_legal[p] = static_cast<int>(n);
    if (result) result[p] = res;
    if (result_ref) result_ref[p] = res_ref;
    if (moves) moves[p] = 0;
    if (res != 0) { ++p; break; }
    SortCanonical(legal, n);
    const chess::Move &m = legal[Pick(seed, game, static_cast<uint64_t>(p), static_cast<uint32_t>(n))];
    if (moves) moves[p] = MoveBits(m);
    b.MakeMove(m);
  }
  return p;
}

// CPU baseline B1: n_threads host threads, each replaying its
// slice of the same deterministic playouts through the reference engine:
// GetGameResult + GetPseudoLegalMoves2 + make/IsKingInCheck/undo + MakeMove.
// Stops after >= min_positions positions in total; returns positions/second.
double ref_bench_playout(const uint8_t *start, uint64_t seed, int n_threads, uint64_t min_positions,
                         int max_plies, uint64_t *positions_out, uint64_t *checksum_out) {
  std::atomic<uint64_t> next_game{0}, positions{0}, checksum{0};
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([&]() {
      chess::Move legal[300];
      uint64_t local_sum = 0;
      while (positions.load(std::memory_order_relaxed) < min_positions) {
        uint64_t game = next_game.fetch_add(1);
        Board b = BoardFromRecord(start);
        uint64_t done = 0;
        for (int p = 0; p < max_plies; ++p) {
          int res_ref = static_cast<int>(b.GetGameResult());
          size_t n = LegalMoves(b, legal);
          int res = CanonicalResult(b, n);
          ++done;
          local_sum += n * 4 + res + (res_ref != res ? 1000003 : 0);
          if (res != 0) break;
          SortCanonical(legal, n);
          b.MakeMove(legal[Pick(seed, game, static_cast<uint64_t>(p), static_cast<uint32_t>(n))]);
        }
        positions.fetch_add(done);
      }
      checksum.fetch_add(local_sum);
    });
  }
  for (auto &x : th) x.join();
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (positions_out) *positions_out = positions.load();
  if (checksum_out) *checksum_out = checksum.load();
  return static_cast<double>(positions.load()) / dt;
}

// ---- the configs[1] workload on the host: n_slots concurrent games resident as chess::Board
// objects, one ply per slot per step, finished slots re-seeded exactly like fpc_playout_step.
// A persistent pool of n_threads workers splits the slots (CPU baseline B1).
struct RefEnv {
  std::vector<std::unique_ptr<Board>> boards;
  std::vector<uint64_t> game;
  std::vector<int> ply;
  std::vector<uint8_t> start;
  uint64_t seed = 0, stride = 0;
  int max_plies = 0, n_threads = 1;
  std::atomic<uint64_t> positions{0}, finished{0}, sum_legal{0};
  // pool
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  uint64_t epoch = 0;
  int pending = 0;
  bool quit = false;

  void StepSlice(int t) {
    chess::Move legal[300];
    const int n = static_cast<int>(boards.size());
    const int lo = static_cast<int>(static_cast<int64_t>(n) * t / n_threads);
    const int hi = static_cast<int>(static_cast<int64_t>(n) * (t + 1) / n_threads);
    uint64_t fin = 0, sl = 0;
    for (int i = lo; i < hi; ++i) {
      Board &b = *boards[i];
      int res_ref = static_cast<int>(b.GetGameResult());
      (void)res_ref;
      size_t k = LegalMoves(b, legal);
      int res = CanonicalResult(b, k);
      sl += k;
      bool done = res != 0;
      if (!done) {
        SortCanonical(legal, k);
        b.MakeMove(legal[Pick(seed, game[i], static_cast<uint64_t>(ply[i]), static_cast<uint32_t>(k))]);
        done = ply[i] + 1 >= max_plies;
      }
      if (done) {
        ++fin;
        boards[i] = std::make_unique<Board>(BoardFromRecord(start.data()));
        game[i] += stride;
        ply[i] = 0;
      } else {
        ++ply[i];
      }
    }
    positions.fetch_add(static_cast<uint64_t>(hi - lo));
    finished.fetch_add(fin);
    sum_legal.fetch_add(sl);
  }

  void Worker(int t) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_go.wait(lk, [&] { return quit || epoch != seen; });
        if (quit) return;
        seen = epoch;
      }
      StepSlice(t);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_one();
      }
    }
  }
};

void *ref_env_create(const uint8_t *start, int n_slots, uint64_t seed, uint64_t first_game, uint64_t stride,
                     int max_plies, int n_threads) {
  RefEnv *e = new RefEnv();
  e->start.assign(start, start + REC);
  e->seed = seed;
  e->stride = stride;
  e->max_plies = max_plies;
  e->n_threads = n_threads < 1 ? 1 : n_threads;
  for (int i = 0; i < n_slots; ++i) {
    e->boards.push_back(std::make_unique<Board>(BoardFromRecord(start)));
    e->game.push_back(first_game + static_cast<uint64_t>(i));
    e->ply.push_back(0);
  }
  for (int t = 0; t < e->n_threads; ++t) e->workers.emplace_back([e, t] { e->Worker(t); });
  return e;
}

void ref_env_step(void *env) {
  RefEnv *e = static_cast<RefEnv *>(env);
  std::unique_lock<std::mutex> lk(e->mu);
  e->pending = e->n_threads;
  ++e->epoch;
  e->cv_go.notify_all();
  e->cv_done.wait(lk, [&] { return e->pending == 0; });
}

void ref_env_stats(void *env, uint64_t *positions, uint64_t *finished, uint64_t *sum_legal) {
  RefEnv *e = static_cast<RefEnv *>(env);
  *positions = e->positions.load();
  *finished = e->finished.load();
  *sum_legal = e->sum_legal.load();
}

void ref_env_get(void *env, int slot, uint8_t *rec, uint64_t *game, int *ply) {
  RefEnv *e = static_cast<RefEnv *>(env);
  RecordFromBoard(*e->boards[slot], rec);
  *game = e->game[slot];
  *ply = e->ply[slot];
}

void ref_env_destroy(void *env) {
  RefEnv *e = static_cast<RefEnv *>(env);
  {
    std::lock_guard<std::mutex> lk(e->mu);
    e->quit = true;
  }
  e->cv_go.notify_all();
  for (auto &w : e->workers) w.join();
  delete e;
}

}  // extern "C"

// pybind11 module `alphazero_cpp`: the reference's Python binding surface
// (/root/reference/src/cpp/wrapper.cpp:15-254) with every rules / encoding operation forwarded to
// the CUDA library through the C-ABI of include/fpc.h.  Same names, argument meaning and error
// behaviour (every failure surfaces as RuntimeError, wrapper.cpp:17-27), so `src/py/mcts.py`,
// `four_player_chess_board.py`, `fen_parser.py` and `alphazero.py` import it unchanged.
//
// Host objects (Board, Move, Node, ...) are thin: a Board is its 208-byte board record, a Move the
// reference's 8-byte move image.  PyTorch (imported as a Python module, not linked) owns device
// memory and streams; the kernels run on torch's current stream and write straight into torch
// tensors.  There is no CPU implementation behind this module: rules calls need a CUDA device.
//
// Geometry: the reference fixes the board size at compile time (engine/board.h:22-24); here it is
// chosen once, before boards are made: environment variable FPC_BOARD_SIZE (14 default, 8 = the
// reference as checked in) or alphazero_cpp.set_board_size(R).
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <array>
#include <map>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "../../include/fpc.h"

namespace py = pybind11;

namespace {

int g_R = 14;
int NSQ() { return g_R * g_R; }
int REC() { return fpc_record_bytes(g_R); }

void check(int rc) {
  if (rc != FPC_OK) throw std::runtime_error(fpc_last_error());
}

// ---- value types (engine/board.h:30-471) ---------------------------------------------------------
enum PieceType : int8_t { PAWN = 0, KNIGHT = 1, BISHOP = 2, ROOK = 3, QUEEN = 4, KING = 5, NO_PIECE = 6 };
enum PlayerColor : int8_t { UNINITIALIZED_PLAYER = -1, RED = 0, BLUE = 1, YELLOW = 2, GREEN = 3 };
enum Team : int8_t { RED_YELLOW = 0, BLUE_GREEN = 1 };
enum GameResult : int8_t { IN_PROGRESS = 0, WIN_RY = 1, WIN_BG = 2, STALEMATE = 3 };

struct Player {
  PlayerColor color = UNINITIALIZED_PLAYER;
  Player() = default;
  explicit Player(PlayerColor c) : color(c) {}
  PlayerColor GetColor() const { return color; }
  Team GetTeam() const { return (color == RED || color == YELLOW) ? RED_YELLOW : BLUE_GREEN; }  // engine/board.h:64-67
  bool operator==(const Player &o) const { return color == o.color; }
  bool operator!=(const Player &o) const { return color != o.color; }
};

struct Piece {
  uint8_t bits = 0x18;  // Piece(false, RED, NO_PIECE), engine/board.h:99-104
  Piece() = default;
  Piece(bool present, PlayerColor c, PieceType t) : bits((uint8_t)((present ? 0x80 : 0) | ((c & 3) << 5) | ((t & 7) << 2))) {}
  Piece(PlayerColor c, PieceType t) : Piece(true, c, t) {}
  Piece(Player p, PieceType t) : Piece(true, p.color, t) {}
  bool Present() const { return bits & 0x80; }
  PlayerColor GetColor() const { return (PlayerColor)((bits >> 5) & 3); }
  PieceType GetPieceType() const { return (PieceType)((bits >> 2) & 7); }
  Player GetPlayer() const { return Player(GetColor()); }
  static std::string ColorToStr(PlayerColor c) {
    switch (c) {
      case RED: return "Red";
      case BLUE: return "Blue";
      case YELLOW: return "Yellow";
      case GREEN: return "Green";
      default: throw std::invalid_argument("Unknown color");
    }
  }
  static std::string PieceTypeToStr(PieceType t) {
    switch (t) {
      case PAWN: return "Pawn";
      case KNIGHT: return "Knight";
      case BISHOP: return "Bishop";
      case ROOK: return "Rook";
      case QUEEN: return "Queen";
      case KING: return "King";
      default: throw std::invalid_argument("Unknown piece type");
    }
  }
  std::string PrettyStr() const {
    if (!Present()) throw std::invalid_argument("Missing piece");
    return ColorToStr(GetColor()) + " " + PieceTypeToStr(GetPieceType());
  }
  bool operator==(const Piece &o) const { return bits == o.bits; }
  bool operator!=(const Piece &o) const { return bits != o.bits; }
};

struct BoardLocation {
  int loc;  // row*R + col, R*R = missing (engine/board.h:190-199)
  BoardLocation() : loc(NSQ()) {}
  BoardLocation(int row, int col) : loc((row < 0 || row >= g_R || col < 0 || col >= g_R) ? NSQ() : row * g_R + col) {}
  static BoardLocation FromSq(int sq) {
    BoardLocation l;
    l.loc = sq;
    return l;
  }
  int GetRow() const { return loc / g_R; }
  int GetCol() const { return loc % g_R; }
  bool operator==(const BoardLocation &o) const { return loc == o.loc; }
  std::string PrettyStr() const {  // engine/board.cpp:1531-1537
    std::string s;
    s += (char)('a' + GetCol());
    s += std::to_string(g_R - GetRow());
    return s + " (" + std::to_string(GetRow()) + ", " + std::to_string(GetCol()) + ")";
  }
};

struct CastlingRights {
  uint8_t bits = 0;  // 0 = absent, else 0x80 | ks<<6 | qs<<5 (engine/board.h:290-291)
  CastlingRights() = default;
  CastlingRights(bool ks, bool qs) : bits((uint8_t)(0x80 | (ks << 6) | (qs << 5))) {}
  bool Kingside() const { return bits & 0x40; }
  bool Queenside() const { return bits & 0x20; }
  bool operator==(const CastlingRights &o) const { return bits == o.bits; }
  bool operator!=(const CastlingRights &o) const { return bits != o.bits; }
};

struct PlacedPiece {
  BoardLocation location;
  Piece piece;
  PlacedPiece() = default;
  PlacedPiece(const BoardLocation &l, const Piece &p) : location(l), piece(p) {}
  std::string PrettyStr() const { return piece.PrettyStr() + " at " + location.PrettyStr(); }  // engine/board.h:463-466
};

// fpchess::Move (src/cpp/move.{h,cpp}) over the 8-byte image of chess::Move (engine/board.h:419-435).
struct Move {
  uint64_t bits;
  static uint64_t pack(int from, int to, uint8_t cap, int promo, int rf, int rt, uint8_t r0, uint8_t r1) {
    return (uint64_t)(uint8_t)from | ((uint64_t)(uint8_t)to << 8) | ((uint64_t)cap << 16) | ((uint64_t)(uint8_t)promo << 24) |
           ((uint64_t)(uint8_t)rf << 32) | ((uint64_t)(uint8_t)rt << 40) | ((uint64_t)r0 << 48) | ((uint64_t)r1 << 56);
  }
  Move() : bits(pack(NSQ(), NSQ(), 0x18, NO_PIECE, NSQ(), NSQ(), 0, 0)) {}
  explicit Move(uint64_t b) : bits(b) {}
  explicit Move(int flat_index) {  // move.cpp:41-61
    bits = fpc_move_from_flat(g_R, flat_index);
    if (bits == ~0ull) throw std::invalid_argument(fpc_last_error());
  }
  Move(int action_plane, const BoardLocation &from) : Move(action_plane * NSQ() + from.loc) {}  // move.cpp:23-39
  Move(const BoardLocation &from, const BoardLocation &to, const Piece &capture, const CastlingRights &initial,
       const CastlingRights &after)
      : bits(pack(from.loc, to.loc, capture.bits, NO_PIECE, NSQ(), NSQ(), initial.bits, after.bits)) {}
  // pawn move: the reference drops the en-passant arguments (engine/board.h:349-359)
  Move(const BoardLocation &from, const BoardLocation &to, const Piece &capture, const BoardLocation &, const Piece &,
       PieceType promotion)
      : bits(pack(from.loc, to.loc, capture.bits, promotion, NSQ(), NSQ(), 0, 0)) {}
  BoardLocation From() const { return BoardLocation::FromSq((int)(bits & 0xff)); }
  BoardLocation To() const { return BoardLocation::FromSq((int)((bits >> 8) & 0xff)); }
  std::tuple<int, int, int> GetIndex() const {  // move.cpp:84-98
    const int flat = fpc_move_flat_index(g_R, bits);
    if (flat < 0)
      throw std::invalid_argument("Invalid move: No corresponding action plane index found. Did you initialize move_index_map?");
    return {flat / NSQ(), From().GetRow(), From().GetCol()};
  }
  int GetFlatIndex() const {  // move.cpp:100-104
    auto [plane, row, col] = GetIndex();
    return plane * NSQ() + row * g_R + col;
  }
};

// ---- device plumbing --------------------------------------------------------------------------------------------
// Rules calls go through a host-buffer context of the C-ABI (fpc_ctx: a stream, device scratch) with page-locked staging
// buffers owned here: one trip = copies in, kernels, copies out, one synchronise -- no Python objects on the way.  PyTorch
// (imported as a Python module, not linked) only appears where the reference returns a tensor.
py::module_ torch() { return py::module_::import("torch"); }

// "cpu" | "gpu" | "cuda" (src/cpp/board.cpp:265-283), extended with "cuda:N" for one process per GPU
struct Device {
  bool cpu_out;
  std::string cuda;  // the device the kernels run on
  int index;         // -1 = the current device
};
Device parse_device(const std::string &d) {
  if (d == "cpu") return {true, "cuda", -1};
  if (d == "gpu" || d == "cuda") return {false, "cuda", -1};
  if (d.rfind("cuda:", 0) == 0) return {false, d, std::atoi(d.c_str() + 5)};
  throw std::invalid_argument("Invalid device argument.");
}
uintptr_t ptr(const py::object &t) { return t.attr("data_ptr")().cast<uintptr_t>(); }

struct Pinned {  // a growing page-locked buffer (fpc_host_alloc)
  void *p = nullptr;
  size_t cap = 0;
  template <class T>
  T *get(size_t count) {
    const size_t bytes = count * sizeof(T);
    if (bytes > cap) {
      if (p) fpc_host_free(p);
      cap = std::max(bytes * 2, (size_t)1 << 20);  // page-locking is slow: few, large steps
      p = fpc_host_alloc(cap);
      if (!p) {
        cap = 0;
        throw std::runtime_error(fpc_last_error());
      }
    }
    return static_cast<T *>(p);
  }
};
constexpr int CTX_CAP = 16384;  // boards per trip; larger batches go in chunks
struct Gpu {
  fpc_ctx *ctx = nullptr;
  Pinned boards, boards2, moves, words, bytes;
};
Gpu &gpu_for(int device) {
  static std::map<std::pair<int, int>, Gpu> all;  // (device, R); lives as long as the process
  if (device < 0) {
    device = fpc_current_device();
    if (device < 0) throw std::runtime_error(std::string("alphazero_cpp (B200 build): ") + fpc_last_error());
  }
  Gpu &g = all[{device, g_R}];
  if (!g.ctx) {
    g.ctx = fpc_ctx_create(device, g_R, CTX_CAP);
    if (!g.ctx) throw std::runtime_error(std::string("alphazero_cpp (B200 build): ") + fpc_last_error());
  }
  return g;
}
// outputs that land in a torch tensor are written on torch's current stream of that device (the allocator's stream)
struct OnTorchStream {
  fpc_ctx *ctx;
  OnTorchStream(fpc_ctx *c, const py::object &tensor) : ctx(c) {
    const uintptr_t st = torch().attr("cuda").attr("current_stream")(tensor.attr("device")).attr("cuda_stream").cast<uintptr_t>();
    fpc_ctx_set_stream(ctx, (void *)st);
  }
  ~OnTorchStream() { fpc_ctx_set_stream(ctx, nullptr); }
};

struct Node;
struct MemoryEntry;

// fpchess::Board (src/cpp/board.{h,cpp}) : chess::Board (engine/board.{h,cpp})
struct Board : std::enable_shared_from_this<Board> {
  std::vector<uint8_t> rec;
  std::shared_ptr<Board> rootState;
  std::shared_ptr<Node> rootNode;
  std::vector<MemoryEntry> memory;
  // what the GPU said about this position (a board only changes through SetTurn): status word, legal moves, attack map.
  // Boards made by TakeAction / ExpandNodes arrive with status and moves from the same trip that made them.
  int obs_status = -1;
  bool obs_moves = false;
  std::vector<uint64_t> legal;
  std::vector<uint8_t> attack;
  void forget() {
    obs_status = -1, obs_moves = false;
    legal.clear(), attack.clear();
  }

  Board() : rec(REC(), 0) {
    std::fill(rec.begin(), rec.begin() + NSQ(), 0x18);
    for (int c = 0; c < 4; ++c) rec[NSQ() + 1 + c] = 0x80, rec[NSQ() + 5 + c] = (uint8_t)NSQ();
  }
  // engine/board.cpp:1172-1248
  Board(Player turn, const py::dict &location_to_piece, const py::object &castling_rights, std::shared_ptr<Board> root) : Board() {
    rec[NSQ()] = (uint8_t)(turn.color & 3);
    for (auto item : location_to_piece) {
      const auto loc = item.first.cast<BoardLocation>();
      const auto piece = item.second.cast<Piece>();
      if (loc.loc >= NSQ()) continue;
      rec[loc.loc] = piece.bits;
      if (piece.Present() && piece.GetPieceType() == KING) rec[NSQ() + 5 + piece.GetColor()] = (uint8_t)loc.loc;
    }
    if (!castling_rights.is_none())
      for (auto item : castling_rights.cast<py::dict>()) {
        py::handle key = item.first;
        const int color = py::isinstance<Player>(key) ? (int)key.cast<Player>().color : (int)key.cast<PlayerColor>();
        rec[NSQ() + 1 + (color & 3)] = item.second.cast<CastlingRights>().bits;
      }
    rootState = std::move(root);
  }
  Player GetTurn() const { return Player((PlayerColor)(rec[NSQ()] & 3)); }
  void SetTurn(const Player &p) {
    rec[NSQ()] = (uint8_t)(p.color & 3);
    forget();
  }
  Piece GetPieceAt(int x, int y) const {
    if (x < 0 || x >= g_R || y < 0 || y >= g_R) throw std::invalid_argument("Index out of bounds");  // engine/board.h:535
    Piece p;
    p.bits = rec[x * g_R + y];
    return p;
  }
  BoardLocation GetBoardLocation(int x, int y) const {
    if (x < 0 || x >= g_R || y < 0 || y >= g_R) throw std::invalid_argument("Index out of bounds");
    return BoardLocation(x, y);
  }
  // piece lists per colour, in the constructor's order K,P,N,B,R,Q (engine/board.cpp:1225-1247); the
  // reference's order afterwards depends on call history and its consumers do not rely on it
  std::vector<std::vector<PlacedPiece>> GetPieces() const {
    static const int order[6] = {KING, PAWN, KNIGHT, BISHOP, ROOK, QUEEN};
    std::vector<std::vector<PlacedPiece>> out(4);
    for (int t : order)
      for (int sq = 0; sq < NSQ(); ++sq) {
        Piece p;
        p.bits = rec[sq];
        if (p.Present() && p.GetPieceType() == t) out[p.GetColor()].emplace_back(BoardLocation::FromSq(sq), p);
      }
    return out;
  }
  std::shared_ptr<Board> GetRootState() { return rootState ? rootState : std::make_shared<Board>(*this); }
  std::vector<MemoryEntry> &GetMemory() { return rootState ? rootState->memory : memory; }

  // rules kernel on this position: status word + legal moves in one trip, kept
  void observe() {
    if (obs_status >= 0 && obs_moves) return;
    Gpu &g = gpu_for(-1);
    uint8_t *in = g.boards.get<uint8_t>(REC());
    memcpy(in, rec.data(), REC());
    uint64_t *mv = g.moves.get<uint64_t>(FPC_MAX_MOVES);
    int32_t *w = g.words.get<int32_t>(2);
    check(fpc_host_observe(g.ctx, in, 1, mv, nullptr, w, w + 1, nullptr, nullptr, -1, nullptr, nullptr));
    obs_status = w[1];
    legal.assign(mv, mv + std::min(std::max(w[0], 0), FPC_MAX_MOVES));
    obs_moves = true;
  }
  // chess::Board::GetGameResult (engine/board.cpp:891-939; order-independent contract)
  GameResult GetGameResult(const std::optional<Player> &) {
    if (obs_status < 0) observe();
    if (obs_status & FPC_STATUS_OVERFLOW) throw std::runtime_error("move buffer overflow");
    return (GameResult)(obs_status & FPC_STATUS_RESULT_MASK);
  }
  // fpchess::Board::GetLegalMoves (src/cpp/board.cpp:94-118), canonical order
  std::vector<std::shared_ptr<Move>> GetLegalMoves() {
    if (!obs_moves) observe();
    std::vector<std::shared_ptr<Move>> out;
    out.reserve(legal.size());
    for (uint64_t m : legal) out.push_back(std::make_shared<Move>(m));
    return out;
  }
  // fpchess::Board::TakeAction (src/cpp/board.cpp:234-239): copy + chess::Board::MakeMove; returns a base Board.
  // One trip makes every child AND observes it (fpc_host_expand), so the children answer GetGameResult / GetLegalMoves
  // from what came back with them.
  static std::vector<std::shared_ptr<Board>> take_actions(const std::vector<std::shared_ptr<Board>> &states, const std::vector<uint64_t> &moves) {
    const size_t total = states.size(), rec_b = (size_t)REC();
    std::vector<std::shared_ptr<Board>> out;
    out.reserve(total);
    Gpu &g = gpu_for(-1);
    for (size_t first = 0; first < total; first += CTX_CAP) {
      const int n = (int)std::min((size_t)CTX_CAP, total - first);
      uint8_t *in = g.boards.get<uint8_t>(n * rec_b), *kids = g.boards2.get<uint8_t>(n * rec_b);
      int32_t *w = g.words.get<int32_t>(3 * (size_t)n);  // err | counts | status
      uint64_t *mv = g.moves.get<uint64_t>((size_t)n);
      for (int i = 0; i < n; ++i) memcpy(in + i * rec_b, states[first + i]->rec.data(), rec_b), mv[i] = moves[first + i];
      check(fpc_host_expand(g.ctx, in, mv, n, kids, w, w + n, w + 2 * n));
      int maxc = 0;
      for (int i = 0; i < n; ++i) {
        if (w[i] != FPC_OK) throw std::runtime_error("piece missing for move");  // engine/board.cpp:1046-1054
        maxc = std::max(maxc, std::min(w[n + i], FPC_MAX_MOVES));
      }
      mv = g.moves.get<uint64_t>((size_t)n * std::max(maxc, 1));
      check(fpc_host_fetch_moves(g.ctx, n, maxc, mv));
      for (int i = 0; i < n; ++i) {
        auto b = std::make_shared<Board>();
        memcpy(b->rec.data(), kids + i * rec_b, rec_b);
        b->obs_status = w[2 * n + i];
        const int c = std::min(std::max(w[n + i], 0), FPC_MAX_MOVES);
        b->legal.assign(mv + (size_t)i * maxc, mv + (size_t)i * maxc + c);
        b->obs_moves = true;
        out.push_back(std::move(b));
      }
    }
    return out;
  }
  std::shared_ptr<Board> TakeAction(const Move &m) { return take_actions({shared_from_this()}, {m.bits})[0]; }
  int CalculateHeuristic(Team team) {  // engine/board.cpp:1263-1292
    Gpu &g = gpu_for(-1);
    uint8_t *in = g.boards.get<uint8_t>(REC());
    memcpy(in, rec.data(), REC());
    in[NSQ()] = (uint8_t)team;  // the kernel evaluates for the team of the side to move
    int32_t *v = g.words.get<int32_t>(1);
    check(fpc_host_heuristic(g.ctx, in, 1, v));
    return v[0];
  }
  // ---- viewer queries (src/cpp/board.cpp:50-57,120-232): one attack-map trip per position, kept ----------------------
  const std::vector<uint8_t> &attack_map() {
    if (attack.empty()) {
      Gpu &g = gpu_for(-1);
      uint8_t *in = g.boards.get<uint8_t>(REC());
      memcpy(in, rec.data(), REC());
      uint8_t *o = g.bytes.get<uint8_t>(NSQ());
      check(fpc_host_attack_maps(g.ctx, in, 1, o));
      attack.assign(o, o + NSQ());
    }
    return attack;
  }
  bool IsAttackedByPlayer(const BoardLocation &l, PlayerColor c) {
    if (l.loc >= NSQ() || c < RED || c > GREEN) return false;  // a missing location has no neighbours inside the box
    return (attack_map()[l.loc] >> (int)c) & 1;
  }
  std::unordered_map<PlayerColor, std::vector<BoardLocation>> GetAttackedSquaresPlayers() {
    const auto &a = attack_map();
    std::unordered_map<PlayerColor, std::vector<BoardLocation>> out;
    for (int c = RED; c <= GREEN; ++c)
      for (int sq = 0; sq < NSQ(); ++sq)
        if ((a[sq] >> c) & 1) out[(PlayerColor)c].push_back(BoardLocation::FromSq(sq));
    return out;
  }
  std::unordered_map<Team, std::vector<BoardLocation>> GetAttackedSquaresTeams() {
    const auto &a = attack_map();
    std::unordered_map<Team, std::vector<BoardLocation>> out;
    for (int t = 0; t < 2; ++t)
      for (int sq = 0; sq < NSQ(); ++sq)
        if ((a[sq] >> (4 + t)) & 1) out[(Team)t].push_back(BoardLocation::FromSq(sq));
    return out;
  }
  std::array<CastlingRights, 4> GetCastlingRights() const {
    std::array<CastlingRights, 4> r;
    for (int c = 0; c < 4; ++c) r[c].bits = rec[NSQ() + 1 + c];
    return r;
  }

  // ---- statics ------------------------------------------------------------------------------------
  static bool IsLegalLocation(int row, int col) {  // engine/board.h:647-654
    const int ia = fpc_invalid_area(g_R);
    if (row < 0 || row >= g_R || col < 0 || col >= g_R) return false;
    const bool corner_col = col < ia || col > g_R - 1 - ia;
    return !((row < ia || row > g_R - 1 - ia) && corner_col);
  }
  static py::object ChangePerspective(const py::object &tensor, int rotation) {  // src/cpp/board.cpp:252-255
    return torch().attr("rot90")(tensor, rotation, py::make_tuple(-2, -1));
  }
  static py::object ParseActionspace(const py::object &flat, const Player &turn) {  // src/cpp/board.cpp:257-263
    py::object v = flat.attr("view")(-1, fpc_num_action_channels(g_R), g_R, g_R);
    return ChangePerspective(v, -(int)turn.color);
  }
  // src/cpp/board.cpp:305-356: [B,24,R,R] f32, the whole batch rotated by the colour of states[0]
  static py::object device_tensor(int n, int channels, const Device &dev) {
    if (!torch().attr("cuda").attr("is_available")().cast<bool>())
      throw std::runtime_error("alphazero_cpp (B200 build): no CUDA device, and there is no CPU fallback");
    return torch().attr("empty")(py::make_tuple(n, channels, g_R, g_R), py::arg("dtype") = torch().attr("float32"), py::arg("device") = dev.cuda);
  }
  static py::object GetEncodedStates(const std::vector<std::shared_ptr<Board>> &states, const std::string &device) {
    const Device dev = parse_device(device);
    const int n = (int)states.size();
    if (n == 0) return torch().attr("zeros")(py::make_tuple(0, FPC_NUM_STATE_CHANNELS, g_R, g_R));
    py::object out = device_tensor(n, FPC_NUM_STATE_CHANNELS, dev);
    Gpu &g = gpu_for(out.attr("device").attr("index").cast<int>());
    OnTorchStream on(g.ctx, out);
    const size_t rec_b = (size_t)REC(), per = (size_t)FPC_NUM_STATE_CHANNELS * NSQ();
    const int k = states[0]->rec[NSQ()] & 3;
    for (int first = 0; first < n; first += CTX_CAP) {
      const int m = std::min(CTX_CAP, n - first);
      uint8_t *in = g.boards.get<uint8_t>(m * rec_b);
      for (int i = 0; i < m; ++i) memcpy(in + i * rec_b, states[first + i]->rec.data(), rec_b);
      check(fpc_host_encode(g.ctx, in, m, k, (float *)ptr(out) + (size_t)first * per));
    }
    return dev.cpu_out ? out.attr("cpu")() : out;
  }
  static py::object GetEncodedState(const Board &state, const std::string &device) {
    return GetEncodedStates({std::make_shared<Board>(state)}, device);
  }
  // FourPlayerChess.get_legal_moves_mask (src/py/four_player_chess_board.py:36-55) in one call: [B,A,R,R] f32
  static py::object LegalMovesMask(const std::vector<std::shared_ptr<Board>> &states, const std::string &device) {
    const Device dev = parse_device(device);
    const int n = (int)states.size(), A = fpc_num_action_channels(g_R);
    if (n == 0) return torch().attr("zeros")(py::make_tuple(0, A, g_R, g_R));
    py::object out = device_tensor(n, A, dev);
    Gpu &g = gpu_for(out.attr("device").attr("index").cast<int>());
    OnTorchStream on(g.ctx, out);
    const size_t rec_b = (size_t)REC(), per = (size_t)A * NSQ();
    for (int first = 0; first < n; first += CTX_CAP) {
      const int m = std::min(CTX_CAP, n - first);
      uint8_t *in = g.boards.get<uint8_t>(m * rec_b);
      for (int i = 0; i < m; ++i) memcpy(in + i * rec_b, states[first + i]->rec.data(), rec_b);
      check(fpc_host_observe(g.ctx, in, m, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, -1, nullptr,
                             (float *)ptr(out) + (size_t)first * per));
    }
    return dev.cpu_out ? out.attr("cpu")() : out;
  }
  // src/cpp/board.cpp:358-422: same result written into the caller's mask tensor
  static py::object GetLegalMovesMask(const std::vector<std::shared_ptr<Board>> &states, const std::string &device, py::object, py::object,
                                      py::object, py::object, py::object legal_moves_masks) {
    py::object m = LegalMovesMask(states, device);
    legal_moves_masks.attr("__getitem__")(py::slice(0, (py::ssize_t)states.size(), 1)).attr("copy_")(m);
    return legal_moves_masks;
  }
  // src/cpp/board.cpp:424-449
  static std::tuple<std::vector<int64_t>, std::vector<int64_t>, std::vector<int64_t>, std::vector<int64_t>> GetLegalMovesIndices(
      const std::vector<std::vector<std::shared_ptr<Move>>> &legal_moves, size_t num_moves) {
    std::vector<int64_t> b(num_moves), p(num_moves), r(num_moves), c(num_moves);
    size_t i = 0;
    for (size_t bi = 0; bi < legal_moves.size(); ++bi)
      for (const auto &m : legal_moves[bi]) {
        if (i >= num_moves) throw std::out_of_range("GetLegalMovesIndices: more moves than num_moves");
        auto [plane, row, col] = m->GetIndex();
        b[i] = (int64_t)bi, p[i] = plane, r[i] = row, c[i] = col;
        ++i;
      }
    return {b, p, r, c};
  }
  std::string Str() const {  // engine/board.cpp:1429-1460 (layout of operator<<)
    static const char *names = "PNBRQKU";
    std::ostringstream os;
    for (int i = 0; i < g_R; ++i) {
      for (int j = 0; j < g_R; ++j) {
        Piece p;
        p.bits = rec[i * g_R + j];
        if (!IsLegalLocation(i, j)) os << "   ";
        else if (!p.Present()) os << " . ";
        else os << (int)p.GetColor() << names[p.GetPieceType()] << ' ';
      }
      os << '\n';
    }
    os << "Turn: Player(" << (int)GetTurn().color << ")\n";
    return os.str();
  }
};

struct MemoryEntry {  // src/cpp/board.h:50-58: a copy of the board + the action-probability tensor
  Board state;
  py::object action;
  MemoryEntry(const Board &s, const py::object &a) : state(s), action(a) {}
};

struct SimpleBoardState {  // engine/board.h:491-497, wrapper.cpp:68-73
  Player turn;
  std::vector<std::vector<PlacedPiece>> pieces;
  std::array<CastlingRights, 4> castlingRights;
  std::unordered_map<PlayerColor, std::vector<BoardLocation>> attackedSquares;
};

struct BoardPool {  // src/cpp/board.h:133-204: the reference's pool never hands out pooled boards either
  explicit BoardPool(size_t) {}
  std::shared_ptr<Board> acquire(const Board &b) { return std::make_shared<Board>(b); }
  void release(std::shared_ptr<Board>) {}
};

// fpchess::Node (src/cpp/node.{h,cpp}).  The per-object tree is host bookkeeping, as in the reference; the
// rules work under it (leaf results, child boards) goes to the GPU in one batch per call.  The
// GPU-resident search is BatchedMCTS / fpc_tree_* (include/fpc.h).
struct Node : std::enable_shared_from_this<Node> {
  int visit_count;
  double C;
  std::shared_ptr<Board> state;
  std::weak_ptr<Node> parent;
  std::shared_ptr<Move> move_made;
  double prior, value_sum = 0;
  std::vector<std::shared_ptr<Node>> children;
  Node(double C_, std::shared_ptr<Board> s, std::shared_ptr<Node> p, std::shared_ptr<Move> m, double prior_, int visits)
      : visit_count(visits), C(C_), state(std::move(s)), parent(p), move_made(std::move(m)), prior(prior_) {}
  bool IsExpanded() const { return !children.empty(); }
  std::shared_ptr<Node> SelectChild() {  // node.cpp:49-78
    int best = -1;
    double best_ucb = -std::numeric_limits<double>::infinity();
    const double lg = std::log(std::sqrt((double)visit_count));
    for (size_t i = 0; i < children.size(); ++i) {
      const auto &c = children[i];
      const double q = c->visit_count > 0 ? c->value_sum / c->visit_count : 0;
      const double ucb = q + C * std::sqrt(lg / (1 + c->visit_count)) * c->prior;
      if (ucb > best_ucb) best = (int)i, best_ucb = ucb;
    }
    if (best < 0) throw std::runtime_error("Failed to select a child.");
    return children[best];
  }
  void Backpropagate(float value) {  // node.cpp:133-142
    value_sum += value;
    visit_count += 1;
    if (auto p = parent.lock()) p->Backpropagate(-value);
  }
  std::shared_ptr<Node> ChooseLeaf() {  // node.cpp:19-47
    auto node = shared_from_this();
    while (node->IsExpanded()) node = node->SelectChild();
    const GameResult r = node->state->GetGameResult(std::nullopt);
    if (r != IN_PROGRESS) {
      node->Backpropagate(r == STALEMATE ? 0.0f : -1.0f);
      return nullptr;
    }
    return node;
  }
  static void BackpropagateNodes(const std::vector<std::shared_ptr<Node>> &nodes, const py::object &values) {  // node.cpp:144-154
    py::list v = values.attr("detach")().attr("to")("cpu").attr("float")().attr("tolist")();
    for (size_t i = 0; i < nodes.size(); ++i) nodes[i]->Backpropagate(v[i].cast<float>());
  }
  // node.cpp:79-131: one child per non-zero policy entry; every child board of the batch is made by ONE
  // make-move launch (the reference copies and makes them one by one)
  static void ExpandNodes(std::vector<std::shared_ptr<Node>> &nodes, const py::object &, const std::vector<std::vector<int64_t>> &nz,
                          const std::vector<double> &values, BoardPool &) {
    std::vector<std::shared_ptr<Board>> parents;
    std::vector<uint64_t> moves;
    std::vector<size_t> owner;
    for (size_t i = 0; i < nz.size(); ++i) {
      const size_t b = (size_t)nz[i][0];
      if (b >= nodes.size()) throw std::out_of_range("ExpandNodes: batch index out of range");
      Move m((int)nz[i][1], BoardLocation((int)nz[i][2], (int)nz[i][3]));
      parents.push_back(nodes[b]->state);
      moves.push_back(m.bits);
      owner.push_back(b);
    }
    if (moves.empty()) return;
    auto boards = Board::take_actions(parents, moves);
    for (size_t i = 0; i < moves.size(); ++i) {
      auto &n = nodes[owner[i]];
      n->children.push_back(std::make_shared<Node>(n->C, boards[i], n, std::make_shared<Move>(moves[i]), values[i], 1));  // node.h:28
    }
  }
};

void set_board_size(py::module_ &m, int R) {
  if (!fpc_supported(R)) throw std::invalid_argument("unsupported board size (14, 13, 10 or 8)");
  g_R = R;
  py::object B = m.attr("Board");
  const int A = fpc_num_action_channels(R);
  B.attr("num_state_channels") = FPC_NUM_STATE_CHANNELS;
  B.attr("state_space_size") = fpc_state_space_size(R);
  B.attr("num_action_channels") = A;
  B.attr("action_space_size") = fpc_action_space_size(R);
  B.attr("action_space_dims") = py::make_tuple(A, R, R);
  B.attr("state_space_dims") = py::make_tuple(FPC_NUM_STATE_CHANNELS, R, R);
  py::object M = m.attr("Move");
  M.attr("num_queen_moves_per_direction") = R - 1;  // move.cpp:18-20
  M.attr("num_queen_moves") = 8 * (R - 1);
  M.attr("num_knight_moves") = 8;
}

}  // namespace

PYBIND11_MODULE(alphazero_cpp, m) {
  m.doc() = "B200-native drop-in for the reference's alphazero_cpp binding (see include/fpc.h)";
  py::register_exception_translator([](std::exception_ptr p) {  // wrapper.cpp:17-27
    try {
      if (p) std::rethrow_exception(p);
    } catch (const py::error_already_set &) {
      throw;
    } catch (const std::exception &e) {
      PyErr_SetString(PyExc_RuntimeError, e.what());
    } catch (...) {
      PyErr_SetString(PyExc_Exception, "An unknown exception occurred.");
    }
  });
  py::module_::import("torch");

  py::enum_<PieceType>(m, "PieceType").value("PAWN", PAWN).value("KNIGHT", KNIGHT).value("BISHOP", BISHOP).value("ROOK", ROOK)
      .value("QUEEN", QUEEN).value("KING", KING).value("NO_PIECE", NO_PIECE).export_values();
  m.def("piece_value", [](PieceType t) { return (int)t; });
  py::enum_<PlayerColor>(m, "PlayerColor").value("UNINITIALIZED_PLAYER", UNINITIALIZED_PLAYER).value("RED", RED).value("BLUE", BLUE)
      .value("YELLOW", YELLOW).value("GREEN", GREEN).export_values();
  m.def("color_value", [](PlayerColor c) { return (int)c; });
  py::enum_<Team>(m, "Team").value("RED_YELLOW", RED_YELLOW).value("BLUE_GREEN", BLUE_GREEN).export_values();
  py::enum_<GameResult>(m, "GameResult").value("IN_PROGRESS", IN_PROGRESS).value("WIN_RY", WIN_RY).value("WIN_BG", WIN_BG)
      .value("STALEMATE", STALEMATE).export_values();

  py::class_<Player>(m, "Player").def(py::init<>()).def(py::init<PlayerColor>()).def("GetColor", &Player::GetColor)
      .def("GetTeam", &Player::GetTeam).def(py::self == py::self).def(py::self != py::self);
  py::class_<Piece>(m, "Piece").def(py::init<>()).def(py::init<bool, PlayerColor, PieceType>()).def(py::init<PlayerColor, PieceType>())
      .def(py::init<Player, PieceType>()).def("GetColor", &Piece::GetColor).def("GetPieceType", &Piece::GetPieceType)
      .def("GetPlayer", &Piece::GetPlayer).def("PieceTypeToStr", [](const Piece &, PieceType t) { return Piece::PieceTypeToStr(t); }, py::arg("type"))
      .def("ColorToStr", [](const Piece &, PlayerColor c) { return Piece::ColorToStr(c); }).def(py::self == py::self).def(py::self != py::self)
      .def("__str__", &Piece::PrettyStr);
  py::class_<BoardLocation>(m, "BoardLocation").def(py::init<>()).def(py::init<int, int>()).def("GetRow", &BoardLocation::GetRow)
      .def("GetCol", &BoardLocation::GetCol).def("__eq__", [](const BoardLocation &a, const BoardLocation &b) { return a == b; })
      .def("__hash__", [](const BoardLocation &l) { return std::hash<int>()(l.GetRow()) ^ std::hash<int>()(l.GetCol()); })
      .def("__str__", &BoardLocation::PrettyStr);
  py::class_<CastlingRights>(m, "CastlingRights").def(py::init<>()).def(py::init<bool, bool>()).def("Kingside", &CastlingRights::Kingside)
      .def("Queenside", &CastlingRights::Queenside).def(py::self == py::self).def(py::self != py::self);
  py::class_<PlacedPiece>(m, "PlacedPiece").def(py::init<>()).def(py::init<const BoardLocation &, const Piece &>())
      .def("GetLocation", [](const PlacedPiece &p) { return p.location; }).def("GetPiece", [](const PlacedPiece &p) { return p.piece; })
      .def("__str__", &PlacedPiece::PrettyStr);

  py::class_<Move, std::shared_ptr<Move>>(m, "Move")
      .def(py::init<>())
      .def(py::init<int, BoardLocation>(), py::arg("action_plane"), py::arg("from"))
      .def(py::init<int>(), py::arg("flat_index"))
      .def(py::init<BoardLocation, BoardLocation, Piece, CastlingRights, CastlingRights>(), py::arg("from"), py::arg("to"),
           py::arg("standard_capture") = Piece(), py::arg("initial_castling_rights") = CastlingRights(), py::arg("castling_rights") = CastlingRights())
      .def(py::init<BoardLocation, BoardLocation, Piece, BoardLocation, Piece, PieceType>(), py::arg("from"), py::arg("to"), py::arg("standard_capture"),
           py::arg("en_passant_location"), py::arg("en_passant_capture"), py::arg("promotion_piece_type") = NO_PIECE)
      .def("From", &Move::From).def("To", &Move::To).def("GetIndex", &Move::GetIndex).def("GetFlatIndex", &Move::GetFlatIndex)
      .def("image", [](const Move &mv) { return mv.bits; }, "the 8-byte chess::Move image (include/fpc.h)");

  py::class_<Board, std::shared_ptr<Board>>(m, "Board")
      .def(py::init<Player, py::dict, py::object, std::shared_ptr<Board>>(), py::arg("turn"), py::arg("location_to_piece"),
           py::arg("castling_rights") = py::none(), py::arg("root_state") = nullptr)
      .def("CalculateHeuristic", &Board::CalculateHeuristic).def("GetTurn", &Board::GetTurn).def("SetTurn", &Board::SetTurn)
      .def_static("GetOpponentValue", [](float v) { return -v; })  // src/cpp/board.cpp:45-48
      .def("GetPieceAt", &Board::GetPieceAt, py::arg("x"), py::arg("y")).def("GetBoardLocation", &Board::GetBoardLocation, py::arg("x"), py::arg("y"))
      .def("GetPieces", &Board::GetPieces)
      .def("GetRootNode", [](Board &b) { return b.rootNode; }).def("SetRootNode", [](Board &b, std::shared_ptr<Node> n) { b.rootNode = std::move(n); })
      .def("GetRootState", &Board::GetRootState).def("SetRootState", [](Board &b, std::shared_ptr<Board> s) { b.rootState = std::move(s); })
      .def("AppendToMemory", [](Board &b, const MemoryEntry &e) { b.GetMemory().push_back(e); })
      .def("GetMemory", [](Board &b) { return b.GetMemory(); })
      .def("GetGameResult", &Board::GetGameResult, py::arg("opt_player") = py::none())
      .def("IsMoveLegal", [](Board &, const Move &) { return false; })  // src/cpp/board.cpp:70-92 always returns false
      // the pygame viewer's queries (src/cpp/board.cpp:50-57,120-232)
      .def("GetSimpleState", [](Board &b) { return SimpleBoardState{b.GetTurn(), b.GetPieces(), b.GetCastlingRights(), b.GetAttackedSquaresPlayers()}; })
      .def("GetAttackedSquaresPlayers", &Board::GetAttackedSquaresPlayers).def("GetAttackedSquaresTeams", &Board::GetAttackedSquaresTeams)
      .def("IsAttackedByPlayer", &Board::IsAttackedByPlayer)
      .def("GetLegalMoves", &Board::GetLegalMoves).def("TakeAction", &Board::TakeAction).def("record", [](const Board &b) { return py::bytes((const char *)b.rec.data(), b.rec.size()); })
      .def_static("ParseActionspace", &Board::ParseActionspace)
      .def_static("IsLegalLocation", [](int r, int c) { return Board::IsLegalLocation(r, c); })
      .def_static("IsLegalLocation", [](const BoardLocation &l) { return l.loc < NSQ() && Board::IsLegalLocation(l.GetRow(), l.GetCol()); })
      .def_static("nRows", [] { return g_R; }).def_static("nCols", [] { return g_R; }).def_static("invalidArea", [] { return fpc_invalid_area(g_R); })
      .def_static("GetOpponent", [](PlayerColor c) { return (PlayerColor)(((int)c + 1) % 4); })  // src/cpp/board.cpp:241-250
      .def_static("GetOpponent", [](const Player &p) { return (PlayerColor)(((int)p.color + 1) % 4); })
      .def_static("ChangePerspective", &Board::ChangePerspective).def_static("GetEncodedState", &Board::GetEncodedState)
      .def_static("GetEncodedStates", &Board::GetEncodedStates).def_static("GetLegalMovesMask", &Board::GetLegalMovesMask)
      .def_static("LegalMovesMask", &Board::LegalMovesMask, py::arg("states"), py::arg("device"))
      .def_static("GetLegalMovesIndices", &Board::GetLegalMovesIndices).def("__str__", &Board::Str);

  py::class_<SimpleBoardState>(m, "SimpleBoardState").def(py::init<>()).def_readwrite("turn", &SimpleBoardState::turn)
      .def_readwrite("pieces", &SimpleBoardState::pieces).def_readwrite("castlingRights", &SimpleBoardState::castlingRights)
      .def_readwrite("attackedSquares", &SimpleBoardState::attackedSquares);
  py::class_<MemoryEntry>(m, "MemoryEntry").def(py::init<const Board &, const py::object &>()).def_readwrite("state", &MemoryEntry::state)
      .def_readwrite("action", &MemoryEntry::action);
  py::class_<BoardPool>(m, "BoardPool").def(py::init<size_t>()).def("acquire", &BoardPool::acquire).def("release", &BoardPool::release);

  py::class_<Node, std::shared_ptr<Node>>(m, "Node")
      .def(py::init<double, std::shared_ptr<Board>, std::shared_ptr<Node>, std::shared_ptr<Move>, double, int>(), py::arg("C"), py::arg("state"),
           py::arg("parent") = nullptr, py::arg("action_taken") = nullptr, py::arg("prior") = 0.0, py::arg("visit_count") = 0)
      .def("GetMoveMade", [](Node &n) { return n.move_made; }).def("GetState", [](Node &n) { return n.state; })
      .def("GetChildren", [](Node &n) { return n.children; }).def("GetVisitCount", [](Node &n) { return n.visit_count; })
      .def("SetVisitCount", [](Node &n, int v) { n.visit_count = v; }).def("IsExpanded", &Node::IsExpanded).def("SelectChild", &Node::SelectChild)
      .def("Backpropagate", &Node::Backpropagate).def_static("BackpropagateNodes", &Node::BackpropagateNodes)
      .def_static("ExpandNodes", &Node::ExpandNodes).def("ChooseLeaf", &Node::ChooseLeaf);

  m.def("set_board_size", [m](int R) mutable { set_board_size(m, R); }, "choose the compiled geometry: 14 (default), 13, 10 or 8");
  m.def("board_size", [] { return g_R; });
  const char *env = std::getenv("FPC_BOARD_SIZE");
  set_board_size(m, env ? std::atoi(env) : 14);
}

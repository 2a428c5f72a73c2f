// Device-side rules of the four-player-chess environment for sm_100a.
//
// Shared pieces of the device-side rules: geometry, piece bytes, the action-plane tables, the compact move,
// the reference's 8-byte move image and make-move on a board record.  One WARP owns one game (fpc_rules.cuh); the
// game's board record (include/fpc.h) is staged in shared memory as a 16x16 "mailbox": cell ((row+1)<<4 | (col+1)),
// every off-board cell and every cut-corner cell holding WALL, so no lookup needs bounds arithmetic.  Rules follow
// the reference engine; each routine cites the reference lines it reproduces (paths relative to
// /root/reference/src/cpp).  No piece list is kept: the reference's piece_list_ order is call-history dependent
// (SURVEY 8a row 9), so lists are produced in canonical order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fpc {

constexpr uint32_t EMPTY = 0x18;  // Piece(false, RED, NO_PIECE), engine/board.h:99
constexpr uint32_t WALL = 0x1C;   // not a reference value: off-board / cut-corner sentinel
constexpr int PAWN = 0, KNIGHT = 1, BISHOP = 2, ROOK = 3, QUEEN = 4, KING = 5, NO_PIECE = 6;
constexpr int NO_SQ = 0xFF;       // mailbox "no square"
constexpr int MAX_MOVES = 300;    // engine/board.h:706
constexpr unsigned FULL = 0xffffffffu;

template <int R_, int IA_>
struct Geo {
  static constexpr int R = R_, IA = IA_, NSQ = R_ * R_;
  static constexpr int REC = ((NSQ + 12 + 15) / 16) * 16;
  static constexpr int OFF_TURN = NSQ, OFF_RIGHTS = NSQ + 1, OFF_KING = NSQ + 5;
  static constexpr int A = 8 * R_ + 8;            // src/cpp/board.cpp:11
  static constexpr int ASZ = A * NSQ;             // action_space_size
  static constexpr int SSZ = 24 * NSQ;            // state_space_size
  static constexpr int MASK_WORDS = (ASZ + 31) / 32;
  static constexpr int PLANE_WORDS = (SSZ + 31) / 32;
  // bit sets are moved 16 bytes at a time: per-game strides are whole uint4s
  static constexpr int MASK_STRIDE = (MASK_WORDS + 3) / 4 * 4, PLANE_STRIDE = (PLANE_WORDS + 3) / 4 * 4;
  static_assert(R_ <= 14, "mailbox is 16x16 with a one-cell border");
  static_assert(ASZ % 4 == 0 && SSZ % 4 == 0, "float4 streaming of the dense outputs");

  // engine/board.h:647-654 IsLegalLocation
  __host__ __device__ static constexpr bool legal(int r, int c) {
    return (unsigned)r < (unsigned)R && (unsigned)c < (unsigned)R &&
           !((r < IA || r > R - 1 - IA) && (c < IA || c > R - 1 - IA));
  }
  __host__ __device__ static constexpr int mb(int r, int c) { return ((r + 1) << 4) | (c + 1); }
  __host__ __device__ static constexpr int mb_of_sq(int sq) { return mb(sq / R, sq % R); }
  __host__ __device__ static constexpr int sq_of_mb(int m) { return ((m >> 4) - 1) * R + (m & 15) - 1; }
};

__device__ __forceinline__ bool present(uint32_t p) { return (p & 0x80u) != 0; }
__device__ __forceinline__ int color_of(uint32_t p) { return (p >> 5) & 3; }
__device__ __forceinline__ int type_of(uint32_t p) { return (p >> 2) & 7; }
__device__ __forceinline__ int team_of(uint32_t p) { return (p >> 5) & 1; }  // engine/board.h:64-67
__device__ __forceinline__ uint32_t mk_piece(int color, int type) { return 0x80u | (color << 5) | (type << 2); }

// The 8 queen directions in the reference's action-plane order (move.cpp:13-14, (dcol,drow)):
// N(0,-1) NW(-1,-1) W(-1,0) SW(-1,1) S(0,1) SE(1,1) E(1,0) NE(1,-1) as mailbox deltas drow*16+dcol.
// Even = orthogonal, odd = diagonal.
__device__ __forceinline__ int qdelta(int dir) {
  return (int)(int8_t)((0xF1011110'0FFFEFF0ull >> (dir * 8)) & 0xff);
}
// Knight jumps in plane order (move.cpp:15-16): (dcol,drow) =
// (-2,-1)(-2,1)(-1,-2)(-1,2)(1,-2)(1,2)(2,-1)(2,1)
__device__ __forceinline__ int kdcol(int k) { return (int)(int8_t)((0x0202'0101'FFFF'FEFEull >> (k * 8)) & 0xff); }
__device__ __forceinline__ int kdrow(int k) { return (int)(int8_t)((0x01FF'02FE'02FE'01FFull >> (k * 8)) & 0xff); }

#ifdef __CUDA_ARCH__
__device__ __forceinline__ int fpc_ffs(uint32_t x) { return __ffs((int)x); }
__device__ __forceinline__ int fpc_clz(uint32_t x) { return __clz((int)x); }
#else
inline int fpc_ffs(uint32_t x) { return __builtin_ffs((int)x); }
inline int fpc_clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
#endif

// engine/board.cpp:23-30 + :1474-1524 GetRookLocationType: 0 kingside, 1 queenside, -1 neither.
template <class G>
__device__ __forceinline__ int rook_location_type(int color, int sq) {
  constexpr int R = G::R, IA = G::IA;
  int ks, qs;
  switch (color) {
    case 0: ks = (R - 1) * R + (R - 4); qs = (R - 1) * R + IA; break;
    case 1: ks = (R - 4) * R; qs = IA * R; break;
    case 2: ks = IA; qs = R - 4; break;
    default: ks = IA * R + (R - 1); qs = (R - 4) * R + (R - 1); break;
  }
  return sq == ks ? 0 : (sq == qs ? 1 : -1);
}

// ---- compact move ------------------------------------------------------------------------
// key = flat*8 + promo (18 bits) in the high bits makes the u32 itself the canonical sort key.
template <class G>
__device__ __forceinline__ uint32_t pack_compact(int from_mb, int to_mb, int plane, int promo, int castle) {
  uint32_t key = (uint32_t)((plane * G::NSQ + G::sq_of_mb(from_mb)) * 8 + promo);
  return (key << 14) | ((uint32_t)castle << 8) | (uint32_t)to_mb;
}
template <class G>
__device__ __forceinline__ void unpack_compact(uint32_t mv, int &from_mb, int &to_mb, int &plane, int &promo,
                                               int &castle) {
  uint32_t key = mv >> 14;
  promo = key & 7;
  uint32_t flat = key >> 3;
  plane = flat / G::NSQ;
  from_mb = G::mb_of_sq(flat - plane * G::NSQ);
  to_mb = mv & 0xff;
  castle = (mv >> 8) & 3;
}

// Rook squares of a castling move: king steps two cells by u = (to-from)/2; the rook stands
// 3 (kingside) / 4 (queenside) cells away and lands on from+u (engine/board.cpp:352-461).
__device__ __forceinline__ void castle_rook(int from_mb, int to_mb, int castle, int &rook_from, int &rook_to) {
  int u = (to_mb - from_mb) / 2;
  rook_to = from_mb + u;
  rook_from = from_mb + u * (castle == 2 ? 3 : 4);  // castle: 1 queenside, 2 kingside
}

// Expand a compact move into the reference's 8-byte image (engine/board.h:419-435).
template <class G>
__device__ uint64_t expand_move(const uint8_t *mb, const uint8_t *rights, uint32_t mv) {
  int from, to, plane, promo, castle;
  unpack_compact<G>(mv, from, to, plane, promo, castle);
  const uint32_t piece = mb[from];
  const int type = type_of(piece), color = color_of(piece);
  const int from_sq = G::sq_of_mb(from), to_sq = G::sq_of_mb(to);
  uint32_t cap = castle ? EMPTY : mb[to];
  uint32_t rf = G::NSQ, rt = G::NSQ, r0 = 0, r1 = 0;
  if (castle) {
    int a, b;
    castle_rook(from, to, castle, a, b);
    rf = G::sq_of_mb(a);
    rt = G::sq_of_mb(b);
  }
  if (type == KING) {  // engine/board.cpp:319-320
    r0 = rights[color];
    r1 = 0x80;
  } else if ((type == ROOK || type == QUEEN) && plane < 8 * (G::R - 1) && !((plane / (G::R - 1)) & 1)) {
    // engine/board.cpp:262-289 (queens reach it through :304-311)
    const int ct = rook_location_type<G>(color, from_sq);
    const uint32_t cur = rights[color];
    const uint32_t ks = (cur >> 6) & 1, qs = (cur >> 5) & 1;
    if (ct == 0 && ks) {
      r0 = cur;
      r1 = 0x80 | (qs << 5);
    } else if (ct == 1 && qs) {
      r0 = cur;
      r1 = 0x80 | (ks << 6);
    }
  }
  return (uint64_t)from_sq | ((uint64_t)to_sq << 8) | ((uint64_t)cap << 16) | ((uint64_t)promo << 24) |
         ((uint64_t)rf << 32) | ((uint64_t)rt << 40) | ((uint64_t)r0 << 48) | ((uint64_t)r1 << 56);
}

// fpchess::Move(int flat_index) (src/cpp/move.cpp:41-61): from = the square, to = from.Relative(...),
// G::NSQ where that leaves the R x R box (BoardLocation's "missing", engine/board.h:194-199).
template <class G>
__device__ __forceinline__ void decode_flat_move(int f, int &from, int &to) {
  constexpr int R = G::R, NSQ = G::NSQ;
  if (f < 0 || f >= G::ASZ) {
    from = to = NSQ;
    return;
  }
  const int type = f / NSQ, pos = f - type * NSQ;
  const int row = pos / R, col = pos - row * R;
  int dr, dc;
  if (type < 8 * (R - 1)) {
    const int dir = type / (R - 1), dist = type - dir * (R - 1) + 1;
    const int d = qdelta(dir);
    const int ur = (d + 24) / 16 - 1, uc = d - ur * 16;
    dr = ur * dist;
    dc = uc * dist;
  } else {
    int k = type - 8 * (R - 1);
    if (k > 7) k = 7;
    dr = kdrow(k);
    dc = kdcol(k);
  }
  const int tr = row + dr, tc = col + dc;
  from = pos;
  to = ((unsigned)tr < (unsigned)R && (unsigned)tc < (unsigned)R) ? tr * R + tc : NSQ;
}

// chess::Board::MakeMove (engine/board.cpp:1028-1096) on the record bytes, for any 8-byte move image
// (index-built moves carry promo = NO_PIECE, rook squares = NSQ, rights-after = 0).  Returns false for
// "piece missing for move" (:1046-1054, thrown after the capture was removed) and off-board squares.
template <class G>
__device__ __forceinline__ bool apply_move_record(uint8_t *b, int from, int to, int promo, int rf, int rt, uint32_t r1) {
  constexpr int NSQ = G::NSQ;
  if (from >= NSQ || to >= NSQ) return false;  // off-board squares: undefined behaviour in the reference
  const int turn = b[G::OFF_TURN] & 3;
  const uint32_t piece = b[from], cap = b[to];
  if (present(cap)) {  // RemovePiece(to), :1040-1044
    b[to] = EMPTY;
    if (type_of(cap) == KING) b[G::OFF_KING + color_of(cap)] = NSQ;
  }
  if (!present(piece)) return false;
  b[from] = EMPTY;
  if (type_of(piece) == KING) b[G::OFF_KING + color_of(piece)] = NSQ;
  const uint32_t placed = promo != NO_PIECE ? mk_piece(turn, promo & 7) : piece;  // :1057-1067
  b[to] = (uint8_t)placed;
  if (type_of(placed) == KING) b[G::OFF_KING + color_of(placed)] = (uint8_t)to;
  if (rf < NSQ && rt < NSQ) {  // :1070-1077
    const uint32_t rook = b[rf];
    b[rf] = EMPTY;
    b[rt] = (uint8_t)rook;
  }
  if (r1 & 0x80) b[G::OFF_RIGHTS + turn] = (uint8_t)r1;  // :1080-1084
  b[G::OFF_TURN] = (uint8_t)((turn + 1) & 3);             // :1088
  return true;
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t game, uint64_t ply) {
  uint64_t z = seed ^ (game * 0x9E3779B97F4A7C15ull) ^ (ply * 0xBF58476D1CE4E5B9ull);
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace fpc

// Device-side rules of the four-player-chess environment for sm_100a.
//
// One WARP owns one game.  The game's board record (include/fpc.h) is staged in shared
// memory as a 16x16 "mailbox": cell ((row+1)<<4 | (col+1)), every off-board cell and every
// cut-corner cell holding WALL, so ray walks need no bounds arithmetic.  Rules follow the
// reference engine; each routine cites the reference lines it reproduces (paths relative to
// /root/reference/src/cpp).  No piece list is kept: the reference's piece_list_ order is
// call-history dependent, so lists are produced in canonical order.
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

// Per-warp shared-memory scratch.
template <class G>
struct alignas(16) WarpScratch {
  uint32_t mask_bits[G::MASK_STRIDE];    // legal-move mask, 1 bit per action
  uint32_t plane_bits[G::PLANE_STRIDE];  // input planes, 1 bit per cell
  uint16_t list[368];             // the ones this call leaves in the dense tensors (fpc_kernels.cu LIST_*)
  uint32_t moves[MAX_MOVES + 4];  // compact moves: key<<14 | castle<<8 | to_mb, key = flat*8 + promo
  uint8_t mb[1024];               // mailbox board in four layouts (see put_cell): rows, columns, diagonals, anti-diagonals
  uint8_t rec[256];               // raw record staging (in and out)
  uint8_t plist[64];              // mailbox squares of the mover's pieces
  uint8_t rights[4];
  uint8_t king[4];                // mailbox square per colour, NO_SQ = captured
  int turn;
  int pad;
};

// A move applied "virtually": a -> EMPTY, c -> EMPTY, b -> bp, d -> dp (order of
// chess::Board::MakeMove, engine/board.cpp:1037-1077).  Unused slots hold 0x1000.
struct Patch {
  int a, b, c, d;
  uint32_t bp, dp;
};

template <bool PATCHED>
__device__ __forceinline__ uint32_t cell(const uint8_t *mb, int i, const Patch &p) {
  uint32_t v = mb[i];
  if (PATCHED) {
    if (i == p.a || i == p.c) v = EMPTY;
    if (i == p.b) v = p.bp;
    if (i == p.d) v = p.dp;
  }
  return v;
}

// ---- the mailbox in four layouts -------------------------------------------------------------------
// Cell (R1, C1) of the 16x16 mailbox (R1 = row+1, C1 = col+1) is stored four times, so that every line
// through a square is 16 contiguous, 16-byte aligned bytes -- ONE shared-memory load per line:
//   rows   mb[        R1 << 4            | C1]   E/W  rays, position C1
//   cols   mb[ 256 | (C1 << 4)           | R1]   S/N  rays, position R1
//   dias   mb[ 512 | ((R1 - C1) & 15) << 4 | C1]   SE/NW rays, position C1 (C1+1 => R1+1)
//   antis  mb[ 768 | ((R1 + C1) & 15) << 4 | C1]   NE/SW rays, position C1 (C1+1 => R1-1)
// A diagonal id wraps round the 16x16 torus, but every wrap crosses the WALL border, so a scan from an
// on-board square never reaches the other part.  Ray walks become bit scans over a line vector: the
// dependent chain of byte loads (one per step) is gone, which matters twice over when the streaming
// expand_kernel keeps the load/store pipe busy (tools/contention.cu).
__host__ __device__ __forceinline__ int idx_col(int m) { return 256 | ((m & 15) << 4) | (m >> 4); }
__host__ __device__ __forceinline__ int idx_dia(int m) { return 512 | ((((m >> 4) - (m & 15)) & 15) << 4) | (m & 15); }
__host__ __device__ __forceinline__ int idx_ant(int m) { return 768 | ((((m >> 4) + (m & 15)) & 15) << 4) | (m & 15); }

__host__ __device__ __forceinline__ void put_cell(uint8_t *mb, int m, uint32_t v) {
  mb[m] = (uint8_t)v;
  mb[idx_col(m)] = (uint8_t)v;
  mb[idx_dia(m)] = (uint8_t)v;
  mb[idx_ant(m)] = (uint8_t)v;
}

#ifdef __CUDA_ARCH__
__device__ __forceinline__ int fpc_ffs(uint32_t x) { return __ffs((int)x); }
__device__ __forceinline__ int fpc_clz(uint32_t x) { return __clz((int)x); }
#else
inline int fpc_ffs(uint32_t x) { return __builtin_ffs((int)x); }
inline int fpc_clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
#endif

struct Line {
  uint32_t w[4];
};
__device__ __forceinline__ Line load_line(const uint8_t *base16) {
  const uint4 v = *reinterpret_cast<const uint4 *>(base16);
  Line l;
  l.w[0] = v.x, l.w[1] = v.y, l.w[2] = v.z, l.w[3] = v.w;
  return l;
}
__device__ __forceinline__ uint32_t line_byte(const Line &l, int pos) {
  const uint32_t w = pos < 8 ? (pos < 4 ? l.w[0] : l.w[1]) : (pos < 12 ? l.w[2] : l.w[3]);
  return (w >> ((pos & 3) * 8)) & 0xffu;
}
__device__ __forceinline__ void line_set(Line &l, int pos, uint32_t val) {
  const uint32_t sh = (pos & 3) * 8, keep = ~(0xffu << sh), ins = val << sh;
  const int wi = pos >> 2;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (wi == i) l.w[i] = (l.w[i] & keep) | ins;
}
// bit i set <=> byte i is not EMPTY.  EMPTY = 0x18; a piece has bit 7, WALL (0x1C) has bit 2.
__device__ __forceinline__ uint32_t line_occupancy(const Line &l) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t t = (l.w[i] | (l.w[i] << 5)) & 0x80808080u;
    m |= ((((t >> 7) * 0x00204081u) >> 21) & 15u) << (4 * i);
  }
  return m;
}
// Nearest non-empty cells on either side of position k (1 <= k <= 14; the WALL border guarantees both).
struct Nearest {
  uint32_t lo, hi;    // piece bytes (WALL for the border)
  bool lo_adj, hi_adj;  // at distance 1
};
__device__ __forceinline__ Nearest line_nearest(const Line &l, int k) {
  const uint32_t occ = line_occupancy(l);
  const int ph = fpc_ffs(occ & (0xfffeu << k)) - 1;
  const int pl = 31 - fpc_clz(occ & ((1u << k) - 1u));
  Nearest n;
  n.hi = line_byte(l, ph);
  n.lo = line_byte(l, pl);
  n.hi_adj = ph == k + 1;
  n.lo_adj = pl == k - 1;
  return n;
}
// Apply one patched cell (mailbox index q -> val) to the four lines through (R1, C1).
__device__ __forceinline__ void patch_lines(Line &row, Line &col, Line &dia, Line &ant, int R1, int C1, int q, uint32_t val) {
  const int qR = q >> 4, qC = q & 15;
  if (qR == R1) line_set(row, qC, val);
  if (qC == C1) line_set(col, qR, val);
  if (((qR - qC) & 15) == ((R1 - C1) & 15)) line_set(dia, qC, val);
  if (((qR + qC) & 15) == ((R1 + C1) & 15)) line_set(ant, qC, val);
}

// chess::Board::GetAttackers2 with limit 1 == IsAttackedByTeam (engine/board.cpp:606-787).
// `s` must be an on-board mailbox square.  Rook rays in the reference run to the edge of the
// R x R box (:632) and bishop rays to the first illegal square (:658); with WALL in the (always
// empty) cut corners both stop at the same attackers.  Sliders (:612-674), pawns (:697-750) and kings
// (:753-772) come out of the four line vectors; knights (:676-694, all eight squares whatever IA is)
// are eight independent byte loads.
template <class G, bool PATCHED>
__device__ bool attacked_by_team(const uint8_t *mb, int team, int s, const Patch &p) {
  const int R1 = s >> 4, C1 = s & 15;
  Line row = load_line(mb + (R1 << 4));
  Line col = load_line(mb + 256 + (C1 << 4));
  Line dia = load_line(mb + 512 + (((R1 - C1) & 15) << 4));
  Line ant = load_line(mb + 768 + (((R1 + C1) & 15) << 4));
  uint32_t kn[8];
  const int r = R1 - 1, c = C1 - 1;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int tr = r + kdrow(k), tc = c + kdcol(k);
    kn[k] = ((unsigned)tr < (unsigned)G::R && (unsigned)tc < (unsigned)G::R) ? cell<PATCHED>(mb, G::mb(tr, tc), p) : WALL;
  }
  if (PATCHED) {
    if (p.a < 256) patch_lines(row, col, dia, ant, R1, C1, p.a, EMPTY);
    if (p.c < 256) patch_lines(row, col, dia, ant, R1, C1, p.c, EMPTY);
    if (p.b < 256) patch_lines(row, col, dia, ant, R1, C1, p.b, p.bp);
    if (p.d < 256) patch_lines(row, col, dia, ant, R1, C1, p.d, p.dp);
  }
  const Nearest nr = line_nearest(row, C1), nc = line_nearest(col, R1), nd = line_nearest(dia, C1), na = line_nearest(ant, C1);
  bool hit = false;
  auto slider = [&](uint32_t v, int straight) {
    const int t = type_of(v);
    hit |= present(v) && team_of(v) == team && (t == QUEEN || t == (straight ? ROOK : BISHOP));
  };
  slider(nr.lo, 1), slider(nr.hi, 1), slider(nc.lo, 1), slider(nc.hi, 1);
  slider(nd.lo, 0), slider(nd.hi, 0), slider(na.lo, 0), slider(na.hi, 0);
  const uint32_t knight = mk_piece(team, KNIGHT) & 0xbfu, king = mk_piece(team, KING) & 0xbfu;  // colour bit 6 ignored: team = bit 5
#pragma unroll
  for (int k = 0; k < 8; ++k) hit |= (kn[k] & 0xbfu) == knight;
  // adjacent kings: the nearest non-empty cell at distance 1 on each of the eight rays
  auto adj_king = [&](uint32_t v, bool adj) { hit |= adj && (v & 0xbfu) == king; };
  adj_king(nr.lo, nr.lo_adj), adj_king(nr.hi, nr.hi_adj), adj_king(nc.lo, nc.lo_adj), adj_king(nc.hi, nc.hi_adj);
  adj_king(nd.lo, nd.lo_adj), adj_king(nd.hi, nd.hi_adj), adj_king(na.lo, na.lo_adj), adj_king(na.hi, na.hi_adj);
  // pawns: RED attacks from the row below, YELLOW from the row above, BLUE from the column to the
  // left, GREEN from the column to the right.  Diagonal neighbours: dia.lo = (R1-1,C1-1), dia.hi =
  // (R1+1,C1+1), ant.lo = (R1+1,C1-1), ant.hi = (R1-1,C1+1).
  const uint32_t d_lo = nd.lo_adj ? nd.lo : EMPTY, d_hi = nd.hi_adj ? nd.hi : EMPTY;
  const uint32_t a_lo = na.lo_adj ? na.lo : EMPTY, a_hi = na.hi_adj ? na.hi : EMPTY;
  if (team == 0) {
    const uint32_t red = mk_piece(0, PAWN), yellow = mk_piece(2, PAWN);
    hit |= a_lo == red || d_hi == red || d_lo == yellow || a_hi == yellow;
  } else {
    const uint32_t blue = mk_piece(1, PAWN), green = mk_piece(3, PAWN);
    hit |= d_lo == blue || a_lo == blue || a_hi == green || d_hi == green;
  }
  return hit;
}

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

// Work item -> moves.  Item (piece, line) of the mover's piece at `from`; a line is one of the four
// lines through the square (0 row: W/E, 1 column: N/S, 2 diagonal: NW/SE, 3 anti-diagonal: SW/NE) and
// yields up to two RUNS of moves, one per direction:
//   kind 0: to = from + delta*(j+1), plane = plane0 + j, j < cnt   (rays, steps, jumps, pushes)
//   kind 1: run 0 only: to = from + delta, plane0, promotion N,B,R,Q for j = 0..3 (engine/board.cpp:82-88)
// Sliders and the king read their line as one 16-byte vector (no per-step loads); a knight item is two of
// the eight jumps; a pawn item is one of push / double push / the two captures.
struct Run {
  int delta, plane0, cnt;
};
// plane-order direction index (move.cpp:13-14) of the "lo" side of a line; the "hi" side is +4
__device__ __forceinline__ int line_lo_dir(int line) { return (0x3102 >> (line * 4)) & 15; }  // W(2) N(0) NW(1) SW(3)

template <class G>
__device__ void gen_item(const uint8_t *mb, int from, int line, Run &lo, Run &hi, int &kind) {
  constexpr int R = G::R;
  const uint32_t p = mb[from];
  const int type = type_of(p), color = color_of(p), team = team_of(p);
  const int R1 = from >> 4, C1 = from & 15;
  const int r = R1 - 1, c = C1 - 1;
  kind = 0;
  lo.delta = lo.plane0 = lo.cnt = 0;
  hi.delta = hi.plane0 = hi.cnt = 0;
  if (type == PAWN) {  // engine/board.cpp:97-177 GetPawnMoves2; `line` = 0 push, 1 double push, 2 / 3 captures
    // forward direction as a plane direction: RED N(0) BLUE E(6) YELLOW S(4) GREEN W(2)
    const int fdir = color == 0 ? 0 : (color == 1 ? 6 : (color == 2 ? 4 : 2));
    const int fwd = qdelta(fdir);
    int to, pdir, dist = 1;
    if (line == 0) {
      to = from + fwd;
      if (mb[to] != EMPTY) return;
      pdir = fdir;
    } else if (line == 1) {
      const bool not_moved = color == 0 ? r == R - 2 : (color == 1 ? c == 1 : (color == 2 ? r == 1 : c == R - 2));
      if (!not_moved || mb[from + fwd] != EMPTY) return;
      to = from + 2 * fwd;
      if (mb[to] != EMPTY) return;
      pdir = fdir;
      dist = 2;
    } else {
      // captures on the two forward diagonals, against the other team only (:153-176)
      // RED: NW(1),NE(7)  YELLOW: SW(3),SE(5)  BLUE: NE(7),SE(5)  GREEN: NW(1),SW(3)
      const int first = line == 2;
      pdir = color == 0 ? (first ? 1 : 7) : (color == 2 ? (first ? 3 : 5) : (color == 1 ? (first ? 7 : 5) : (first ? 1 : 3)));
      to = from + qdelta(pdir);
      const uint32_t o = mb[to];
      if (!present(o) || team_of(o) == team) return;
    }
    lo.delta = to - from;
    lo.plane0 = pdir * (R - 1) + dist - 1;
    // promotion line (:58-76): RED row R/4, YELLOW row 3R/4, BLUE col 3R/4, GREEN col R/4
    const int tr = (to >> 4) - 1, tc = (to & 15) - 1;
    const bool promo = color == 0 ? tr == R / 4 : (color == 2 ? tr == 3 * R / 4 : (color == 1 ? tc == 3 * R / 4 : tc == R / 4));
    kind = promo ? 1 : 0;
    lo.cnt = promo ? 4 : 1;
    return;
  }
  if (type == KNIGHT) {  // engine/board.cpp:179-207: |drow| runs 1..IA-1 only; jumps 2*line and 2*line+1
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 2 * line + h;
      const int dr = kdrow(k), dc = kdcol(k);
      Run &out = h ? hi : lo;
      if ((dr < 0 ? -dr : dr) >= G::IA) continue;
      const int tr = r + dr, tc = c + dc;
      if ((unsigned)tr >= (unsigned)R || (unsigned)tc >= (unsigned)R) continue;
      const int to = G::mb(tr, tc);
      const uint32_t o = mb[to];
      if (o == WALL || (present(o) && team_of(o) == team)) continue;
      out.delta = to - from;
      out.plane0 = 8 * (R - 1) + k;
      out.cnt = 1;
    }
    return;
  }
  const bool straight = line < 2;
  if (type == BISHOP ? straight : (type == ROOK ? !straight : (type != QUEEN && type != KING))) return;
  // the line through `from`: sliders (engine/board.cpp:209-238 AddMovesFromIncrMovement2) and king steps (:322-341)
  const int k = line == 1 ? R1 : C1;
  const Line l = load_line(line == 0 ? mb + (R1 << 4)
                                     : (line == 1 ? mb + 256 + (C1 << 4)
                                                  : (line == 2 ? mb + 512 + (((R1 - C1) & 15) << 4) : mb + 768 + (((R1 + C1) & 15) << 4))));
  const int dlo = line_lo_dir(line), dhi = dlo + 4;
  lo.delta = qdelta(dlo);
  lo.plane0 = dlo * (R - 1);
  hi.delta = qdelta(dhi);
  hi.plane0 = dhi * (R - 1);
  if (type == KING) {
    const uint32_t a = line_byte(l, k - 1), b = line_byte(l, k + 1);
    lo.cnt = !(a == WALL || (present(a) && team_of(a) == team));
    hi.cnt = !(b == WALL || (present(b) && team_of(b) == team));
    return;
  }
  const uint32_t occ = line_occupancy(l);
  const int ph = fpc_ffs(occ & (0xfffeu << k)) - 1;
  const int pl = 31 - fpc_clz(occ & ((1u << k) - 1u));
  const uint32_t bh = line_byte(l, ph), bl = line_byte(l, pl);
  hi.cnt = ph - k - 1 + (present(bh) && team_of(bh) != team);
  lo.cnt = k - pl - 1 + (present(bl) && team_of(bl) != team);
}

// Castling candidate (engine/board.cpp:343-465).  side: 0 queenside, 1 kingside.  Returns the
// compact move or 0.  The rook may belong to the partner (:437 compares teams).
template <class G>
__device__ uint32_t gen_castle(const uint8_t *mb, int from, int color, uint32_t rights, int side) {
  constexpr int R = G::R;
  const bool allowed = side ? (rights >> 6) & 1 : (rights >> 5) & 1;
  if (!allowed) return 0;
  // unit step king -> rook: RED ks E(6)/qs W(2), BLUE ks S(4)/qs N(0), YELLOW ks W/qs E, GREEN ks N/qs S
  const int udir = color == 0 ? (side ? 6 : 2) : (color == 1 ? (side ? 4 : 0) : (color == 2 ? (side ? 2 : 6) : (side ? 0 : 4)));
  const int u = qdelta(udir);
  const int nb = side ? 2 : 3;
  const int r = (from >> 4) - 1, c = (from & 15) - 1;
  const int ur = (u + 24) / 16 - 1;  // row component of u in {-1,0,1}
  const int uc = u - ur * 16;
  const int rr = r + ur * (nb + 1), rc = c + uc * (nb + 1);
  if ((unsigned)rr >= (unsigned)R || (unsigned)rc >= (unsigned)R) return 0;  // Relative() -> missing (engine/board.h:194-199)
  const uint32_t rook = mb[from + u * (nb + 1)];
  if (!present(rook) || type_of(rook) != ROOK || team_of(rook) != (color & 1)) return 0;
  for (int k = 1; k <= nb; ++k)
    if (mb[from + u * k] != EMPTY) return 0;
  Patch none{0x1000, 0x1000, 0x1000, 0x1000, 0, 0};
  const int other = 1 - (color & 1);
  if (attacked_by_team<G, false>(mb, other, from + u, none)) return 0;  // :456
  if (attacked_by_team<G, false>(mb, other, from, none)) return 0;
  return pack_compact<G>(from, from + 2 * u, udir * (R - 1) + 1, NO_PIECE, side ? 2 : 1);
}

// src/cpp/board.cpp:59-68 IsKingSafeAfterMove as an attack test on the virtually patched board.
template <class G>
__device__ bool king_safe_after(const uint8_t *mb, const uint8_t *king, int turn, uint32_t mv) {
  int from, to, plane, promo, castle;
  unpack_compact<G>(mv, from, to, plane, promo, castle);
  const uint32_t piece = mb[from];
  Patch p{from, to, 0x1000, 0x1000, piece, 0};
  if (castle) {
    castle_rook(from, to, castle, p.c, p.d);
    p.dp = mb[p.c];
  }
  const int ks = type_of(piece) == KING ? to : king[turn];
  if (ks == NO_SQ) return true;  // engine/board.cpp:945-948
  return !attacked_by_team<G, true>(mb, 1 - (turn & 1), ks, p);
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

// chess::Board::MakeMove (engine/board.cpp:1028-1096) on the mailbox, for a generator move.
template <class G>
__device__ __forceinline__ void make_compact(WarpScratch<G> &s, uint32_t mv) {
  int from, to, plane, promo, castle;
  unpack_compact<G>(mv, from, to, plane, promo, castle);
  const int turn = s.turn;
  const uint32_t piece = s.mb[from], cap = s.mb[to];
  const int type = type_of(piece);
  // rights carried by the move (see expand_move)
  if (type == KING) {
    s.rights[turn] = 0x80;
  } else if ((type == ROOK || type == QUEEN) && plane < 8 * (G::R - 1) && !((plane / (G::R - 1)) & 1)) {
    const int ct = rook_location_type<G>(turn, G::sq_of_mb(from));
    const uint32_t cur = s.rights[turn];
    if (ct == 0 && ((cur >> 6) & 1)) s.rights[turn] = 0x80 | (cur & 0x20);
    else if (ct == 1 && ((cur >> 5) & 1)) s.rights[turn] = 0x80 | (cur & 0x40);
  }
  if (present(cap) && type_of(cap) == KING) s.king[color_of(cap)] = NO_SQ;
  put_cell(s.mb, from, EMPTY);
  put_cell(s.mb, to, promo != NO_PIECE ? mk_piece(turn, promo) : piece);
  if (type == KING) s.king[turn] = to;
  if (castle) {
    int rf, rt;
    castle_rook(from, to, castle, rf, rt);
    const uint32_t rook = s.mb[rf];
    put_cell(s.mb, rf, EMPTY);
    put_cell(s.mb, rt, rook);
  }
  s.turn = (turn + 1) & 3;
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

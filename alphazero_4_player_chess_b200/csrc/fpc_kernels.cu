// Kernels and C-ABI of the batched four-player-chess environment (sm_100a).
// See include/fpc.h for the boundary and fpc_device.cuh for the rules.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/fpc.h"
#include "fpc_rules.cuh"

namespace fpc {

int fail(int code, const std::string &msg);
int cuda_check(cudaError_t e, const char *what);

constexpr int WARPS_PER_BLOCK = 4;  // one warp per game
constexpr int BLOCK_THREADS = WARPS_PER_BLOCK * 32;

// Dense outputs.  The reference tensors are dense f32 ([n,24,R,R] planes, [n,8R+8,R,R] mask) holding
// ~40 and ~19 ones per game among 4,704 and 23,520 cells; writing them is the HBM-bound part of the
// path (113 KB per game at 14x14).  rules_kernel produces them as bit sets (1 bit per cell, 3.5 KB
// per game, zero-filled and set straight in global memory = L2); expand_kernel streams the bits out as f32.
// The two kernels run on different streams so that the expansion of one batch overlaps the integer work
// of the next (FPC_FLAG_ASYNC_DENSE, fpc_join).

static_assert(STATUS_IN_CHECK == FPC_STATUS_IN_CHECK && STATUS_CAN_TAKE_KING == FPC_STATUS_CAN_TAKE_KING &&
                  STATUS_OVERFLOW == FPC_STATUS_OVERFLOW && STATUS_FINISHED == FPC_STATUS_FINISHED &&
                  STATUS_CHECK == FPC_STATUS_CHECK,
              "status bits of fpc_rules.cuh and include/fpc.h");

#define CK(expr)                                   \
  do {                                             \
    int rc_ = cuda_check((expr), #expr);           \
    if (rc_ != FPC_OK) return rc_;                 \
  } while (0)

// The rules kernel: one warp per game (fpc_rules.cuh).  Legal moves, result, the bit sets of the dense outputs,
// and (playout) the move choice and make-move.  It touches only the board store and compact per-game outputs.
template <class G>
__global__ void __launch_bounds__(BLOCK_THREADS) rules_kernel(const __grid_constant__ ObserveParams P) {
  __shared__ RulesScratch<G> scratch[WARPS_PER_BLOCK];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * WARPS_PER_BLOCK + wib;
  if (g >= P.n) return;
  rules_warp<G>(P, scratch[wib], g, lane);
}

// bits -> dense f32 (0.0 / 1.0).  One tensor per launch half: games x (words_per_game words ->
// floats_per_game floats).  A warp takes 32 words of one game per iteration with one coalesced
// load, hands them round with shuffles and issues 8 store instructions of 512 contiguous bytes.
// Pure streaming: no shared memory, one pass, HBM-write bound.
constexpr int EXPAND_THREADS = 128;
constexpr int EXPAND_ITERS = 1;  // 32-word groups per warp (finer CTAs stream better: tools/sweep.sh)

struct ExpandHalf {
  const uint32_t *bits;  // [n][words]
  float *out;            // [n][floats]
  int words, stride, floats, groups;  // per game: words used, words allocated, floats, ceil(words / 32)
};

__global__ void __launch_bounds__(512)
    expand_kernel(const __grid_constant__ ExpandHalf A, const __grid_constant__ ExpandHalf B, int n, int iters) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = (unsigned long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const unsigned long long total_a = (unsigned long long)n * A.groups, total_b = (unsigned long long)n * B.groups;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    unsigned long long grp = warp * iters + it;
    const ExpandHalf *H = &A;
    if (grp >= total_a) {
      grp -= total_a;
      if (grp >= total_b) return;
      H = &B;
    }
    const unsigned long long game = grp / H->groups;
    const int w0 = (int)(grp - game * H->groups) * 32;
    const uint32_t mine = w0 + lane < H->words ? __ldg(H->bits + game * H->stride + w0 + lane) : 0u;
    float4 *dst = reinterpret_cast<float4 *>(H->out + game * H->floats);
    // almost every word is zero (~40 ones among 4,704 cells, ~19 among 23,520): a store whose four
    // words are all zero needs no shuffle and no bit arithmetic -- warp-uniform test on one ballot
    const unsigned nz = __ballot_sync(FULL, mine != 0u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // store j, lane l: float4 number w0*8 + j*32 + l of this game = word w0 + j*4 + l/8, nibble l%8
      const int f4 = w0 * 8 + j * 32 + lane;
      float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if ((nz >> (j * 4)) & 15u) {
        const uint32_t w = __shfl_sync(FULL, mine, j * 4 + (lane >> 3));
        const uint32_t nib = (w >> ((lane & 7) * 4)) & 15u;
        v.x = (nib & 1u) ? 1.0f : 0.0f;
        v.y = (nib & 2u) ? 1.0f : 0.0f;
        v.z = (nib & 4u) ? 1.0f : 0.0f;
        v.w = (nib & 8u) ? 1.0f : 0.0f;
      }
      if (f4 * 4 < H->floats) __stcs(dst + f4, v);
    }
  }
}

// chess::Board::MakeMove (engine/board.cpp:1028-1096) for arbitrary 8-byte moves, or for
// index-built moves (src/cpp/move.cpp:41-61) when flat != null.  Works on the record bytes.
template <class G>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    make_kernel(const uint8_t *in, const uint64_t *moves, const int32_t *flat, int n, uint8_t *out, int32_t *err) {
  __shared__ alignas(16) uint8_t recs[WARPS_PER_BLOCK][256];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * WARPS_PER_BLOCK + wib;
  if (g >= n) return;
  uint8_t *b = recs[wib];
  if (lane < G::REC / 16)
    reinterpret_cast<uint4 *>(b)[lane] = reinterpret_cast<const uint4 *>(in + (size_t)g * G::REC)[lane];
  __syncwarp();
  if (lane == 0) {
    constexpr int NSQ = G::NSQ;
    int from, to, promo = NO_PIECE, rf = NSQ, rt = NSQ;
    uint32_t r1 = 0;
    if (flat) {
      decode_flat_move<G>(flat[g], from, to);
    } else {
      const uint64_t m = moves[g];
      from = (int)(m & 0xff);
      to = (int)((m >> 8) & 0xff);
      promo = (int)((m >> 24) & 0xff);
      rf = (int)((m >> 32) & 0xff);
      rt = (int)((m >> 40) & 0xff);
      r1 = (uint32_t)((m >> 56) & 0xff);
    }
    const int code = apply_move_record<G>(b, from, to, promo, rf, rt, r1) ? FPC_OK : FPC_ERR_MOVE;
    if (err) err[g] = code;
  }
  __syncwarp();
  if (lane < G::REC / 16)
    reinterpret_cast<uint4 *>(out + (size_t)g * G::REC)[lane] = reinterpret_cast<const uint4 *>(b)[lane];
}

// chess::Board::CalculateHeuristic (engine/board.cpp:1263-1292) for the team to move.
template <class G>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) heuristic_kernel(const uint8_t *in, int n, int32_t *value) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * WARPS_PER_BLOCK + wib;
  if (g >= n) return;
  const uint8_t *b = in + (size_t)g * G::REC;
  const int team = b[G::OFF_TURN] & 1;
  int h = 0;
  for (int sq = lane; sq < G::NSQ; sq += 32) {
    const uint32_t p = b[sq];
    if (present(p) && type_of(p) != KING) {
      const int t = type_of(p);
      const int v = t == PAWN ? 1 : (t == ROOK ? 5 : (t == QUEEN ? 9 : 3));
      h += team_of(p) == team ? v : -v;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(FULL, h, o);
  if (lane == 0) value[g] = h;
}

// ---- host side ----------------------------------------------------------------------------

static thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
int cuda_check(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return FPC_OK;
  return fail(FPC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// Per host thread and device: the stream expand_kernel runs on, its events, and the double-buffered
// bit-set workspace between rules_kernel and expand_kernel.
struct SideState {
  cudaStream_t side = nullptr;  // expand_kernel: lowest priority
  cudaStream_t hi = nullptr;    // rules_kernel when dense outputs are wanted: highest priority, so that its
                                // CTAs are dispatched ahead of the remaining CTAs of a running expansion
  cudaEvent_t fork = nullptr;
  cudaEvent_t rules_done[2] = {nullptr, nullptr}, expand_done[2] = {nullptr, nullptr};
  bool expand_recorded[2] = {false, false};
  uint32_t *bits[2] = {nullptr, nullptr};
  size_t bits_words = 0;
  int parity = 0;
  int last = -1;  // buffer of the most recent expansion (fpc_join)
  // Dense tensors whose content is known exactly: the list of ones the latest call left in them
  // (FPC_FLAG_INCREMENTAL updates such tensors in place instead of rewriting them).
  struct Tracked {
    const float *planes = nullptr, *mask = nullptr;
    int n = 0, R = 0;
    uint16_t *lists = nullptr;
    size_t lists_games = 0;
    unsigned long long stamp = 0;
  };
  Tracked tracked[4];
  unsigned long long stamp = 0;
  // optional CUDA-event timing of expand_kernel on its own stream (fpc_profile_enable / _read)
  bool prof_on = false;
  int prof_n = 0;
  std::vector<cudaEvent_t> prof_ev;   // expand_kernel: start/stop pairs
  std::vector<cudaEvent_t> prof_ev_r; // rules_kernel (dense path): start/stop pairs
};
constexpr int PROF_MAX = 4096;
static thread_local SideState g_side[16];

static int side_state(SideState **out, size_t words, cudaStream_t st) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(FPC_ERR_ARG, "device ordinal out of range");
  SideState &S = g_side[dev];
  if (!S.side) {
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CK(cudaStreamCreateWithPriority(&S.side, cudaStreamNonBlocking, least));
    CK(cudaStreamCreateWithPriority(&S.hi, cudaStreamNonBlocking, greatest));
    CK(cudaEventCreateWithFlags(&S.fork, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&S.rules_done[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&S.expand_done[i], cudaEventDisableTiming));
    }
  }
  if (words > S.bits_words) {
    // growing the workspace: earlier launches may still use the old one
    if (S.bits[0]) {
      CK(cudaStreamSynchronize(st));
      CK(cudaStreamSynchronize(S.side));
      cudaFree(S.bits[0]);
      cudaFree(S.bits[1]);
      S.bits[0] = S.bits[1] = nullptr;
      S.bits_words = 0;
    }
    CK(cudaMalloc(&S.bits[0], words * sizeof(uint32_t)));
    CK(cudaMalloc(&S.bits[1], words * sizeof(uint32_t)));
    S.bits_words = words;
  }
  *out = &S;
  return FPC_OK;
}

struct DenseOut {
  float *planes, *mask;
  int flags;
};

// The record slot of a (planes, mask, n, R) output set: an existing one (content known: *known = true) or the least
// recently used one, re-targeted (content unknown until a full dense call has run).
static int tracked_slot(SideState *S, const DenseOut &d, int n, int R, cudaStream_t st, SideState::Tracked **out, bool *known) {
  SideState::Tracked *hit = nullptr, *lru = &S->tracked[0];
  for (auto &t : S->tracked) {
    if (t.lists && t.planes == d.planes && t.mask == d.mask && t.n == n && t.R == R) hit = &t;
    if (t.stamp < lru->stamp) lru = &t;
  }
  *known = hit != nullptr;
  SideState::Tracked *t = hit ? hit : lru;
  if (!hit) {
    if (t->lists_games < (size_t)n) {
      if (t->lists) {
        CK(cudaStreamSynchronize(st));
        CK(cudaStreamSynchronize(S->side));
        CK(cudaStreamSynchronize(S->hi));
        cudaFree(t->lists);
        t->lists = nullptr;
      }
      CK(cudaMalloc(&t->lists, (size_t)n * LIST_STRIDE * sizeof(uint16_t)));
      t->lists_games = (size_t)n;
    }
    t->planes = d.planes, t->mask = d.mask, t->n = n, t->R = R;
  }
  t->stamp = ++S->stamp;
  *out = t;
  return FPC_OK;
}

// rules_kernel on the caller's stream; expand_kernel on the side stream once the rules kernel has
// written the bit sets.  Unless FPC_FLAG_ASYNC_DENSE is set the caller's stream then waits for the
// expansion.  after_rules (may be a no-op) runs right after the rules kernel is enqueued: the
// host-buffer entry points start their device-to-host copies of the compact results there.
template <class G, class F>
static int launch_observe(ObserveParams p, DenseOut d, cudaStream_t st, F after_rules) {
  if (p.n == 0) return FPC_OK;
  const bool dense = d.planes || d.mask;
  SideState *S = nullptr;
  int b = 0;
  if (dense) {
    if ((reinterpret_cast<uintptr_t>(d.planes) | reinterpret_cast<uintptr_t>(d.mask)) & 15)
      return fail(FPC_ERR_ARG, "planes / mask must be 16-byte aligned");
    const size_t pw = (size_t)p.n * G::PLANE_STRIDE, mw = (size_t)p.n * G::MASK_STRIDE;
    int rc = side_state(&S, pw + mw, st);
    if (rc != FPC_OK) return rc;
    SideState::Tracked *trk = nullptr;
    bool known = false;
    rc = tracked_slot(S, d, p.n, G::R, st, &trk, &known);
    if (rc != FPC_OK) return rc;
    p.lists = trk->lists;
    p.list_cells = d.planes != nullptr;
    p.list_flats = d.mask != nullptr;
    if ((d.flags & FPC_FLAG_INCREMENTAL) && known) {
      // the tensors hold exactly the recorded ones: clear those, set the new ones, no expansion.  A still running
      // expansion into the same tensors (an earlier FPC_FLAG_ASYNC_DENSE call) must finish first.
      if (S->last >= 0) CK(cudaStreamWaitEvent(st, S->expand_done[S->last], 0));
      p.inc_planes = d.planes;
      p.inc_mask = d.mask;
      const int blocks_inc = (p.n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
      rules_kernel<G><<<blocks_inc, BLOCK_THREADS, 0, st>>>(p);
      CK(cudaGetLastError());
      return after_rules();
    }
    b = S->parity;
    S->parity ^= 1;
    p.plane_bits = d.planes ? S->bits[b] : nullptr;
    p.mask_bits = d.mask ? S->bits[b] + pw : nullptr;
  }
  const int blocks = (p.n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  if (!dense) {
    rules_kernel<G><<<blocks, BLOCK_THREADS, 0, st>>>(p);
    CK(cudaGetLastError());
    return after_rules();
  }
  // fork: caller's stream -> high-priority stream (rules) -> back to the caller's stream; the
  // expansion that last read this bit buffer must be done before the rules kernel rewrites it
  CK(cudaEventRecord(S->fork, st));
  CK(cudaStreamWaitEvent(S->hi, S->fork, 0));
  if (S->expand_recorded[b]) CK(cudaStreamWaitEvent(S->hi, S->expand_done[b], 0));
  // Both kernels ask for the same shared-memory carveout: CTAs of rules_kernel (24 KB of shared memory each) and of
  // expand_kernel (none) share SMs, and an SM only changes its L1 / shared split when it is empty.
  {
    static thread_local bool carveout_set[16] = {false};
    int dev = 0;
    CK(cudaGetDevice(&dev));
    const char *env = getenv("FPC_CARVEOUT");
    const int pct = env ? atoi(env) : 100;
    if (dev >= 0 && dev < 16 && !carveout_set[dev] && pct >= 0) {
      CK(cudaFuncSetAttribute(rules_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
      CK(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
      carveout_set[dev] = true;
    }
  }
  const bool prof_r = S->prof_on && S->prof_n < PROF_MAX;
  if (prof_r) CK(cudaEventRecord(S->prof_ev_r[2 * S->prof_n], S->hi));
  rules_kernel<G><<<blocks, BLOCK_THREADS, 0, S->hi>>>(p);
  CK(cudaGetLastError());
  if (prof_r) CK(cudaEventRecord(S->prof_ev_r[2 * S->prof_n + 1], S->hi));
  CK(cudaEventRecord(S->rules_done[b], S->hi));
  CK(cudaStreamWaitEvent(st, S->rules_done[b], 0));
  int rc = after_rules();
  if (rc != FPC_OK) return rc;
  {
    CK(cudaStreamWaitEvent(S->side, S->rules_done[b], 0));
    ExpandHalf A{p.plane_bits, d.planes, G::PLANE_WORDS, G::PLANE_STRIDE, G::SSZ, d.planes ? (G::PLANE_WORDS + 31) / 32 : 0};
    ExpandHalf B{p.mask_bits, d.mask, G::MASK_WORDS, G::MASK_STRIDE, G::ASZ, d.mask ? (G::MASK_WORDS + 31) / 32 : 0};
    const unsigned long long groups = (unsigned long long)p.n * (A.groups + B.groups);
    static const int iters = getenv("FPC_EXPAND_ITERS") ? atoi(getenv("FPC_EXPAND_ITERS")) : EXPAND_ITERS;
    static const int threads = getenv("FPC_EXPAND_THREADS") ? atoi(getenv("FPC_EXPAND_THREADS")) : EXPAND_THREADS;
    const unsigned long long warps = (groups + iters - 1) / iters;
    const unsigned long long grid = (warps + threads / 32 - 1) / (threads / 32);
    const bool prof = S->prof_on && S->prof_n < PROF_MAX;
    if (prof) CK(cudaEventRecord(S->prof_ev[2 * S->prof_n], S->side));
    expand_kernel<<<(unsigned)grid, threads, 0, S->side>>>(A, B, p.n, iters);
    CK(cudaGetLastError());
    if (prof) CK(cudaEventRecord(S->prof_ev[2 * S->prof_n++ + 1], S->side));
    CK(cudaEventRecord(S->expand_done[b], S->side));
    S->expand_recorded[b] = true;
    S->last = b;
    if (!(d.flags & FPC_FLAG_ASYNC_DENSE)) CK(cudaStreamWaitEvent(st, S->expand_done[b], 0));
  }
  return FPC_OK;
}
template <class G>
static int launch_make(const uint8_t *in, const uint64_t *moves, const int32_t *flat, int n, uint8_t *out,
                       int32_t *err, cudaStream_t st) {
  if (n == 0) return FPC_OK;
  const int blocks = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  make_kernel<G><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(in, moves, flat, n, out, err);
  return cuda_check(cudaGetLastError(), "make_kernel launch");
}
template <class G>
static int launch_heuristic(const uint8_t *in, int n, int32_t *v, cudaStream_t st) {
  if (n == 0) return FPC_OK;
  const int blocks = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  heuristic_kernel<G><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(in, n, v);
  return cuda_check(cudaGetLastError(), "heuristic_kernel launch");
}

#define FPC_DISPATCH(R, CALL)                                              \
  switch (R) {                                                             \
    case 14: { using G = Geo<14, 3>; return CALL; }                        \
    case 13: { using G = Geo<13, 3>; return CALL; }                        \
    case 10: { using G = Geo<10, 2>; return CALL; }                        \
    case 8: { using G = Geo<8, 2>; return CALL; }                          \
    default: return fail(FPC_ERR_ARG, "unsupported board size R=" + std::to_string(R)); \
  }

static int ia_of(int R) { return R == 14 || R == 13 ? 3 : (R == 10 || R == 8 ? 2 : -1); }

template <class F>
static int do_observe(int R, const ObserveParams &p, DenseOut d, cudaStream_t st, F after_rules) {
  FPC_DISPATCH(R, (launch_observe<G>(p, d, st, after_rules)));
}
static int do_observe(int R, const ObserveParams &p, DenseOut d, cudaStream_t st) {
  return do_observe(R, p, d, st, [] { return FPC_OK; });
}
static int do_make(int R, const uint8_t *in, const uint64_t *moves, const int32_t *flat, int n, uint8_t *out,
                   int32_t *err, cudaStream_t st) {
  FPC_DISPATCH(R, launch_make<G>(in, moves, flat, n, out, err, st));
}
static int do_heuristic(int R, const uint8_t *in, int n, int32_t *v, cudaStream_t st) {
  FPC_DISPATCH(R, launch_heuristic<G>(in, n, v, st));
}

}  // namespace fpc

using namespace fpc;

extern "C" {

const char *fpc_last_error(void) { return g_err.c_str(); }
int fpc_version(void) { return 100; }
int fpc_supported(int R) { return ia_of(R) > 0; }
int fpc_invalid_area(int R) { return ia_of(R); }
int fpc_record_bytes(int R) { return fpc_supported(R) ? ((R * R + 12 + 15) / 16) * 16 : FPC_ERR_ARG; }
int fpc_num_action_channels(int R) { return fpc_supported(R) ? 8 * R + 8 : FPC_ERR_ARG; }
int fpc_action_space_size(int R) { return fpc_supported(R) ? (8 * R + 8) * R * R : FPC_ERR_ARG; }
int fpc_state_space_size(int R) { return fpc_supported(R) ? 24 * R * R : FPC_ERR_ARG; }

static const int kQd[8][2] = {{0, -1}, {-1, -1}, {-1, 0}, {-1, 1}, {0, 1}, {1, 1}, {1, 0}, {1, -1}};   // move.cpp:13-14
static const int kKd[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}}; // move.cpp:15-16

uint64_t fpc_move_from_flat(int R, int flat) {
  const int nsq = R * R;
  if (!fpc_supported(R) || flat < 0 || flat >= (8 * R + 8) * nsq) {
    fail(FPC_ERR_ARG, "flat index out of range");
    return ~0ull;
  }
  const int type = flat / nsq, pos = flat % nsq, row = pos / R, col = pos % R;
  int dr, dc;
  if (type < 8 * (R - 1)) {
    const int dir = type / (R - 1), dist = type % (R - 1) + 1;
    dc = kQd[dir][0] * dist;
    dr = kQd[dir][1] * dist;
  } else {
    int k = type - 8 * (R - 1);
    if (k > 7) k = 7;
    dc = kKd[k][0];
    dr = kKd[k][1];
  }
  const int tr = row + dr, tc = col + dc;
  const uint64_t to = (tr < 0 || tr >= R || tc < 0 || tc >= R) ? nsq : tr * R + tc;
  return (uint64_t)pos | (to << 8) | (0x18ull << 16) | (6ull << 24) | ((uint64_t)nsq << 32) | ((uint64_t)nsq << 40);
}

int fpc_move_flat_index(int R, uint64_t move) {
  if (!fpc_supported(R)) return fail(FPC_ERR_ARG, "unsupported board size");
  const int from = (int)(move & 0xff), to = (int)((move >> 8) & 0xff);
  // BoardLocation::GetRow/GetCol (engine/board.h:203-204) also decode the "missing" value R*R
  const int dx = to % R - from % R, dy = to / R - from / R;
  int plane = -1;
  for (int i = 0; i < 8 && plane < 0; ++i)
    for (int d = 1; d <= R - 1; ++d)
      if (dx == kQd[i][0] * d && dy == kQd[i][1] * d) { plane = i * (R - 1) + d - 1; break; }
  for (int i = 0; i < 8 && plane < 0; ++i)
    if (dx == kKd[i][0] && dy == kKd[i][1]) plane = 8 * (R - 1) + i;
  if (plane < 0) return -1;
  return plane * R * R + (from / R) * R + from % R;
}

int fpc_record_from_fen(int R, const char *fen, int honour_castling, uint8_t *h_record) {
  if (!fpc_supported(R) || !fen || !h_record) return fail(FPC_ERR_ARG, "fpc_record_from_fen: bad argument");
  const int nsq = R * R, rec = fpc_record_bytes(R);
  std::string text;
  for (const char *c = fen; *c; ++c)
    if (*c != '\n' && *c != '\r' && *c != ' ') text += *c;
  auto split = [](const std::string &str, char sep) {
    std::vector<std::string> out;
    size_t at = 0;
    for (;;) {
      const size_t next = str.find(sep, at);
      out.push_back(str.substr(at, next == std::string::npos ? std::string::npos : next - at));
      if (next == std::string::npos) break;
      at = next + 1;
    }
    return out;
  };
  const std::vector<std::string> parts = split(text, '-');
  if (parts.size() < 5) return fail(FPC_ERR_ARG, "FEN string has too few fields");
  static const char turns[] = "RBYG";
  const char *tp = parts[0].size() == 1 ? strchr(turns, parts[0][0]) : nullptr;
  if (!tp || !*tp) return fail(FPC_ERR_ARG, "Invalid player character in FEN string");
  memset(h_record, 0, rec);
  memset(h_record, 0x18, nsq);
  h_record[nsq] = (uint8_t)(tp - turns);
  for (int c = 0; c < 4; ++c) h_record[nsq + 1 + c] = 0x80, h_record[nsq + 5 + c] = (uint8_t)nsq;
  const std::vector<std::string> ks = split(parts[2], ','), qs = split(parts[3], ',');
  if (ks.size() != 4) return fail(FPC_ERR_ARG, "Invalid kingside castling availability in FEN string");
  if (qs.size() != 4) return fail(FPC_ERR_ARG, "Invalid queenside castling availability in FEN string");
  if (honour_castling)
    for (int c = 0; c < 4; ++c) h_record[nsq + 1 + c] = (uint8_t)(0x80 | ((ks[c] == "1") << 6) | ((qs[c] == "1") << 5));
  const std::vector<std::string> rows = split(parts.back(), '/');
  if ((int)rows.size() > R) return fail(FPC_ERR_ARG, "Too many rows in piece placement");
  static const char colors[] = "rbyg", types[] = "PNBRQK";
  for (int row = 0; row < (int)rows.size(); ++row) {
    int col = 0;
    for (const std::string &cell : split(rows[row], ',')) {
      if (cell.empty()) return fail(FPC_ERR_ARG, "Empty column string in piece placement");
      const char *cp = strchr(colors, cell[0]);
      if (cp && *cp) {
        const char *ty = cell.size() == 2 ? strchr(types, cell[1]) : nullptr;
        if (!ty || !*ty) return fail(FPC_ERR_ARG, "Piece placement string for player must be of length 2");
        if (col >= R) return fail(FPC_ERR_ARG, "Piece placement outside the board");
        const int color = (int)(cp - colors), type = (int)(ty - types);
        h_record[row * R + col] = (uint8_t)(0x80 | (color << 5) | (type << 2));
        if (type == 5) h_record[nsq + 5 + color] = (uint8_t)(row * R + col);
        ++col;
      } else if (cell == "x") {
        ++col;
      } else {
        char *end = nullptr;
        const long n = strtol(cell.c_str(), &end, 10);
        if (!end || *end || n <= 0) return fail(FPC_ERR_ARG, "Invalid number of empty spaces in piece placement");
        col += (int)n;
      }
    }
  }
  return FPC_OK;
}

int fpc_profile_enable(int on) {
  SideState *S = nullptr;
  int rc = side_state(&S, 0, nullptr);
  if (rc != FPC_OK) return rc;
  if (on && S->prof_ev.empty()) {
    S->prof_ev.resize(2 * PROF_MAX);
    for (auto &e : S->prof_ev) CK(cudaEventCreate(&e));
    S->prof_ev_r.resize(2 * PROF_MAX);
    for (auto &e : S->prof_ev_r) CK(cudaEventCreate(&e));
  }
  S->prof_on = on != 0;
  S->prof_n = 0;
  return FPC_OK;
}

int fpc_profile_read(int *launches, double *expand_ms, double *rules_ms) {
  SideState *S = nullptr;
  int rc = side_state(&S, 0, nullptr);
  if (rc != FPC_OK) return rc;
  CK(cudaStreamSynchronize(S->side));
  CK(cudaStreamSynchronize(S->hi));
  double total = 0, total_r = 0;
  for (int i = 0; i < S->prof_n; ++i) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, S->prof_ev[2 * i], S->prof_ev[2 * i + 1]));
    total += ms;
    CK(cudaEventElapsedTime(&ms, S->prof_ev_r[2 * i], S->prof_ev_r[2 * i + 1]));
    total_r += ms;
  }
  if (launches) *launches = S->prof_n;
  if (expand_ms) *expand_ms = total;
  if (rules_ms) *rules_ms = total_r;
  return FPC_OK;
}

int fpc_join(void *stream) {
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(FPC_ERR_ARG, "device ordinal out of range");
  SideState &S = g_side[dev];
  if (S.last >= 0) CK(cudaStreamWaitEvent((cudaStream_t)stream, S.expand_done[S.last], 0));
  return FPC_OK;
}

int fpc_observe(int R, const uint8_t *d_boards, int n, uint64_t *d_moves, int32_t *d_flat, int32_t *d_counts,
                int32_t *d_status, float *d_planes, const int32_t *d_k, int k_all, float *d_mask, int flags,
                void *stream) {
  if (n < 0 || (n > 0 && !d_boards)) return fail(FPC_ERR_ARG, "fpc_observe: bad boards/n");
  ObserveParams p{};
  p.boards_in = d_boards;
  p.n = n;
  p.need_movegen = (d_moves || d_flat || d_counts || d_status || d_mask) ? 1 : 0;
  p.moves = d_moves;
  p.flat = d_flat;
  p.counts = d_counts;
  p.status = d_status;
  p.k = d_k;
  p.k_all = k_all;
  return do_observe(R, p, DenseOut{d_planes, d_mask, flags}, (cudaStream_t)stream);
}

int fpc_encode(int R, const uint8_t *d_boards, int n, const int32_t *d_k, int k_all, float *d_planes, int flags,
               void *stream) {
  if (!d_planes && n > 0) return fail(FPC_ERR_ARG, "fpc_encode: null output");
  return fpc_observe(R, d_boards, n, nullptr, nullptr, nullptr, nullptr, d_planes, d_k, k_all, nullptr, flags,
                     stream);
}

int fpc_make_moves(int R, const uint8_t *d_in, const uint64_t *d_moves, int n, uint8_t *d_out, int32_t *d_err,
                   void *stream) {
  if (n < 0 || (n > 0 && (!d_in || !d_moves || !d_out))) return fail(FPC_ERR_ARG, "fpc_make_moves: bad argument");
  return do_make(R, d_in, d_moves, nullptr, n, d_out, d_err, (cudaStream_t)stream);
}

int fpc_make_index(int R, const uint8_t *d_in, const int32_t *d_flat, int n, uint8_t *d_out, int32_t *d_err,
                   void *stream) {
  if (n < 0 || (n > 0 && (!d_in || !d_flat || !d_out))) return fail(FPC_ERR_ARG, "fpc_make_index: bad argument");
  return do_make(R, d_in, nullptr, d_flat, n, d_out, d_err, (cudaStream_t)stream);
}

int fpc_heuristic(int R, const uint8_t *d_boards, int n, int32_t *d_value, void *stream) {
  if (n < 0 || (n > 0 && (!d_boards || !d_value))) return fail(FPC_ERR_ARG, "fpc_heuristic: bad argument");
  return do_heuristic(R, d_boards, n, d_value, (cudaStream_t)stream);
}

static int playout_params(ObserveParams &p, uint8_t *d_boards, int n, uint64_t seed, uint64_t *d_game, int32_t *d_ply,
                          const uint8_t *d_start, int max_plies, uint64_t game_stride, uint64_t *d_chosen,
                          int32_t *d_counts, int32_t *d_status, const int32_t *d_k, int k_all,
                          uint64_t *d_counters) {
  if (n < 0 || (n > 0 && (!d_boards || !d_game || !d_ply || !d_start)) || max_plies <= 0)
    return fail(FPC_ERR_ARG, "fpc_playout_step: bad argument");
  p = ObserveParams{};
  p.boards_in = d_boards;
  p.boards_out = d_boards;
  p.n = n;
  p.need_movegen = 1;
  p.counts = d_counts;
  p.status = d_status;
  p.k = d_k;
  p.k_all = k_all;
  p.playout = 1;
  p.seed = seed;
  p.game = d_game;
  p.ply = d_ply;
  p.start = d_start;
  p.max_plies = max_plies;
  p.game_stride = game_stride;
  p.chosen = d_chosen;
  p.counters = reinterpret_cast<unsigned long long *>(d_counters);
  return FPC_OK;
}

int fpc_playout_step(int R, uint8_t *d_boards, int n, uint64_t seed, uint64_t *d_game, int32_t *d_ply,
                     const uint8_t *d_start, int max_plies, uint64_t game_stride, uint64_t *d_chosen,
                     int32_t *d_counts, int32_t *d_status, float *d_planes, const int32_t *d_k, int k_all,
                     float *d_mask, uint64_t *d_counters, int flags, void *stream) {
  ObserveParams p;
  int rc = playout_params(p, d_boards, n, seed, d_game, d_ply, d_start, max_plies, game_stride, d_chosen, d_counts,
                          d_status, d_k, k_all, d_counters);
  if (rc != FPC_OK) return rc;
  return do_observe(R, p, DenseOut{d_planes, d_mask, flags}, (cudaStream_t)stream);
}

// ---- host-buffer context ---------------------------------------------------------------------

struct fpc_ctx {
  int device, R, max_n, rec;
  cudaStream_t stream;
  uint8_t *d_boards, *d_boards2, *d_start;
  uint64_t *d_moves, *d_game;
  int32_t *d_flat, *d_counts, *d_status, *d_ply, *d_err;
  float *d_planes, *d_mask;   // lazily allocated for host-destination dense outputs
  uint8_t h_start_cache[256];
  bool start_valid;
};

fpc_ctx *fpc_ctx_create(int device, int R, int max_n) {
  if (!fpc_supported(R) || max_n <= 0) {
    fail(FPC_ERR_ARG, "fpc_ctx_create: bad argument");
    return nullptr;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
    fail(FPC_ERR_CUDA, "fpc_ctx_create: no usable CUDA device (there is no CPU fallback)");
    return nullptr;
  }
  if (cuda_check(cudaSetDevice(device), "cudaSetDevice") != FPC_OK) return nullptr;
  fpc_ctx *c = new fpc_ctx();
  memset(c, 0, sizeof *c);
  c->device = device;
  c->R = R;
  c->max_n = max_n;
  c->rec = fpc_record_bytes(R);
  const size_t n = (size_t)max_n;
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_boards, n * c->rec) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_boards2, n * c->rec) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_start, c->rec) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_moves, n * FPC_MAX_MOVES * sizeof(uint64_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_flat, n * FPC_MAX_MOVES * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_game, n * sizeof(uint64_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_counts, n * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_status, n * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_ply, n * sizeof(int32_t)) == cudaSuccess;
  ok = ok && cudaMalloc(&c->d_err, n * sizeof(int32_t)) == cudaSuccess;
  if (!ok) {
    fail(FPC_ERR_CUDA, std::string("fpc_ctx_create: ") + cudaGetErrorString(cudaGetLastError()));
    fpc_ctx_destroy(c);
    return nullptr;
  }
  return c;
}

void fpc_ctx_destroy(fpc_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaFree(c->d_boards);
  cudaFree(c->d_boards2);
  cudaFree(c->d_start);
  cudaFree(c->d_moves);
  cudaFree(c->d_flat);
  cudaFree(c->d_game);
  cudaFree(c->d_counts);
  cudaFree(c->d_status);
  cudaFree(c->d_ply);
  cudaFree(c->d_err);
  cudaFree(c->d_planes);
  cudaFree(c->d_mask);
  delete c;
}

void *fpc_ctx_stream(fpc_ctx *c) { return c ? (void *)c->stream : nullptr; }


static int ctx_check(fpc_ctx *c, int n, const char *who) {
  if (!c) return fail(FPC_ERR_ARG, std::string(who) + ": null context");
  if (n < 0 || n > c->max_n) return fail(FPC_ERR_ARG, std::string(who) + ": n exceeds the context capacity");
  return cuda_check(cudaSetDevice(c->device), "cudaSetDevice");
}

int fpc_ctx_sync(fpc_ctx *c) {
  int rc = ctx_check(c, 0, "fpc_ctx_sync");
  if (rc != FPC_OK) return rc;
  rc = fpc_join(c->stream);
  if (rc != FPC_OK) return rc;
  return cuda_check(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
}

int fpc_host_observe(fpc_ctx *c, const uint8_t *h_boards, int n, uint64_t *h_moves, int32_t *h_flat,
                     int32_t *h_counts, int32_t *h_status, float *h_planes, float *d_planes, int k_all,
                     float *h_mask, float *d_mask) {
  int rc = ctx_check(c, n, "fpc_host_observe");
  if (rc != FPC_OK) return rc;
  if (n == 0) return FPC_OK;
  if (!h_boards) return fail(FPC_ERR_ARG, "fpc_host_observe: null boards");
  const size_t N = (size_t)n;
  const size_t ssz = (size_t)fpc_state_space_size(c->R), asz = (size_t)fpc_action_space_size(c->R);
  if (h_planes && !d_planes) {
    if (!c->d_planes) CK(cudaMalloc(&c->d_planes, (size_t)c->max_n * ssz * sizeof(float)));
    d_planes = c->d_planes;
  }
  if (h_mask && !d_mask) {
    if (!c->d_mask) CK(cudaMalloc(&c->d_mask, (size_t)c->max_n * asz * sizeof(float)));
    d_mask = c->d_mask;
  }
  CK(cudaMemcpyAsync(c->d_boards, h_boards, N * c->rec, cudaMemcpyHostToDevice, c->stream));
  rc = fpc_observe(c->R, c->d_boards, n, h_moves ? c->d_moves : nullptr, h_flat ? c->d_flat : nullptr,
                   (h_counts || h_moves || h_flat) ? c->d_counts : nullptr, h_status ? c->d_status : nullptr,
                   d_planes, nullptr, k_all, d_mask, 0, c->stream);
  if (rc != FPC_OK) return rc;
  if (h_counts) CK(cudaMemcpyAsync(h_counts, c->d_counts, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (h_status) CK(cudaMemcpyAsync(h_status, c->d_status, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (h_moves)
    CK(cudaMemcpyAsync(h_moves, c->d_moves, N * FPC_MAX_MOVES * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
  if (h_flat)
    CK(cudaMemcpyAsync(h_flat, c->d_flat, N * FPC_MAX_MOVES * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (h_planes) CK(cudaMemcpyAsync(h_planes, d_planes, N * ssz * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (h_mask) CK(cudaMemcpyAsync(h_mask, d_mask, N * asz * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FPC_OK;
}

static int host_make(fpc_ctx *c, const uint8_t *h_in, const uint64_t *h_moves, const int32_t *h_flat, int n,
                     uint8_t *h_out, int32_t *h_err) {
  int rc = ctx_check(c, n, "fpc_host_make");
  if (rc != FPC_OK) return rc;
  if (n == 0) return FPC_OK;
  if (!h_in || !h_out || (!h_moves && !h_flat)) return fail(FPC_ERR_ARG, "fpc_host_make: null argument");
  const size_t N = (size_t)n;
  CK(cudaMemcpyAsync(c->d_boards, h_in, N * c->rec, cudaMemcpyHostToDevice, c->stream));
  if (h_moves) {
    CK(cudaMemcpyAsync(c->d_moves, h_moves, N * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    rc = fpc_make_moves(c->R, c->d_boards, c->d_moves, n, c->d_boards2, c->d_err, c->stream);
  } else {
    CK(cudaMemcpyAsync(c->d_flat, h_flat, N * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    rc = fpc_make_index(c->R, c->d_boards, c->d_flat, n, c->d_boards2, c->d_err, c->stream);
  }
  if (rc != FPC_OK) return rc;
  CK(cudaMemcpyAsync(h_out, c->d_boards2, N * c->rec, cudaMemcpyDeviceToHost, c->stream));
  if (h_err) CK(cudaMemcpyAsync(h_err, c->d_err, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FPC_OK;
}

int fpc_host_make_moves(fpc_ctx *c, const uint8_t *h_in, const uint64_t *h_moves, int n, uint8_t *h_out,
                        int32_t *h_err) {
  return host_make(c, h_in, h_moves, nullptr, n, h_out, h_err);
}
int fpc_host_make_index(fpc_ctx *c, const uint8_t *h_in, const int32_t *h_flat, int n, uint8_t *h_out,
                        int32_t *h_err) {
  return host_make(c, h_in, nullptr, h_flat, n, h_out, h_err);
}

// Device alias of a pinned (page-locked, mapped) host pointer, or null for pageable memory.
static void *mapped_alias(const void *h) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, h) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int fpc_host_playout_step(fpc_ctx *c, uint8_t *h_boards, int n, uint64_t seed, uint64_t *h_game, int32_t *h_ply,
                          const uint8_t *h_start, int max_plies, uint64_t game_stride, int32_t *h_counts,
                          int32_t *h_status, float *d_planes, int k_all, float *d_mask, int flags) {
  int rc = ctx_check(c, n, "fpc_host_playout_step");
  if (rc != FPC_OK) return rc;
  if (n == 0) return FPC_OK;
  if (!h_boards || !h_game || !h_ply || !h_start) return fail(FPC_ERR_ARG, "fpc_host_playout_step: null argument");
  const size_t N = (size_t)n;
  // the start record only changes between runs: upload it when it differs from the cached copy
  if (!c->start_valid || memcmp(c->h_start_cache, h_start, c->rec) != 0) {
    memcpy(c->h_start_cache, h_start, c->rec);
    CK(cudaMemcpyAsync(c->d_start, c->h_start_cache, c->rec, cudaMemcpyHostToDevice, c->stream));
    c->start_valid = true;
  }
  ObserveParams p;
  // Zero-copy: when every host buffer is pinned, the rules kernel reads the records straight from
  // host memory and writes its results straight back over PCIe -- no staging copies, one launch.
  uint8_t *m_boards = (uint8_t *)mapped_alias(h_boards);
  uint64_t *m_game = (uint64_t *)mapped_alias(h_game);
  int32_t *m_ply = (int32_t *)mapped_alias(h_ply);
  int32_t *m_counts = h_counts ? (int32_t *)mapped_alias(h_counts) : nullptr;
  int32_t *m_status = h_status ? (int32_t *)mapped_alias(h_status) : nullptr;
  if (m_boards && m_game && m_ply && (!h_counts || m_counts) && (!h_status || m_status)) {
    rc = playout_params(p, m_boards, n, seed, m_game, m_ply, c->d_start, max_plies, game_stride, nullptr, m_counts,
                        m_status, nullptr, k_all, nullptr);
    if (rc != FPC_OK) return rc;
    rc = do_observe(c->R, p, DenseOut{d_planes, d_mask, flags}, c->stream);
    if (rc != FPC_OK) return rc;
    CK(cudaStreamSynchronize(c->stream));
    return FPC_OK;
  }
  CK(cudaMemcpyAsync(c->d_boards, h_boards, N * c->rec, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_game, h_game, N * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_ply, h_ply, N * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  rc = playout_params(p, c->d_boards, n, seed, c->d_game, c->d_ply, c->d_start, max_plies, game_stride, nullptr,
                      c->d_counts, c->d_status, nullptr, k_all, nullptr);
  if (rc != FPC_OK) return rc;
  // the compact results go back to the host while the dense outputs are still being written
  rc = do_observe(c->R, p, DenseOut{d_planes, d_mask, flags}, c->stream, [&]() -> int {
    CK(cudaMemcpyAsync(h_boards, c->d_boards, N * c->rec, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_game, c->d_game, N * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_ply, c->d_ply, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (h_counts) CK(cudaMemcpyAsync(h_counts, c->d_counts, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (h_status) CK(cudaMemcpyAsync(h_status, c->d_status, N * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    return FPC_OK;
  });
  if (rc != FPC_OK) return rc;
  CK(cudaStreamSynchronize(c->stream));
  return FPC_OK;
}

}  // extern "C"

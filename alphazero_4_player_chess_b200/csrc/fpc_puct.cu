// Batched PUCT (select / expand / backup) for sm_100a: one warp owns one game's tree in HBM.
// Boundary and array layout: include/fpc.h (struct fpc_tree).  Semantics follow the reference's
// fpchess::Node (src/cpp/node.{h,cpp}) driven by src/py/mcts.py; each routine cites the lines.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <string>

#include "../../include/fpc.h"
#include "fpc_device.cuh"

namespace fpc {

int fail(int code, const std::string &msg);  // fpc_kernels.cu
int cuda_check(cudaError_t e, const char *what);

#define CK(expr)                                   \
  do {                                             \
    int rc_ = cuda_check((expr), #expr);           \
    if (rc_ != FPC_OK) return rc_;                 \
  } while (0)

constexpr int SEL_WARPS = 4;
constexpr int ERR_NODE_CAP = 1, ERR_BOARD_CAP = 2, ERR_NO_CHILD = 4, ERR_MOVE = 8;

template <class G>
__global__ void __launch_bounds__(256) tree_reset_kernel(const fpc_tree T, const uint8_t *roots) {
  const int g = blockIdx.x;
  const size_t slab = (size_t)g * T.node_cap;
  if (threadIdx.x == 0) {
    T.parent[slab] = -1;
    T.first_child[slab] = 0;
    T.n_children[slab] = 0;
    T.visits[slab] = 1;  // mcts.py:30
    T.move_flat[slab] = -1;
    T.board_idx[slab] = 0;
    T.value_sum[slab] = 0.0;
    T.prior[slab] = 0.0f;
    T.n_nodes[g] = 1;
    T.n_boards[g] = 1;
    T.leaf[g] = -1;
    T.dropped[g] = 0;
    T.error[g] = 0;
  }
  for (int i = threadIdx.x; i < G::REC; i += blockDim.x)
    T.boards[(size_t)g * T.board_cap * G::REC + i] = roots[(size_t)g * G::REC + i];
}

// Node::ChooseLeaf's descent (node.cpp:21-26) + Node::SelectChild (node.cpp:49-78).
template <class G>
__global__ void __launch_bounds__(SEL_WARPS * 32) tree_descend_kernel(const fpc_tree T) {
  __shared__ alignas(16) uint8_t recs[SEL_WARPS][256];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = blockIdx.x * SEL_WARPS + wib;
  if (g >= T.n_games) return;
  if (T.dropped[g]) {
    if (lane == 0) T.leaf[g] = -1;
    return;
  }
  const size_t slab = (size_t)g * T.node_cap;
  const int32_t *visits = T.visits + slab, *first_child = T.first_child + slab, *n_children = T.n_children + slab;
  const double *value_sum = T.value_sum + slab;
  const float *prior = T.prior + slab;
  // One dependent memory round per level: while scanning the children of `node` every lane also fetches
  // its child's own child range, and the winner's (first_child, n_children, visits) travel with the arg-max.
  int node = 0;
  int nc = n_children[0], fc = first_child[0], nv = visits[0];
  while (nc > 0) {
    const double lg = log(sqrt((double)nv));
    double best = -CUDART_INF;
    int best_i = 0x7fffffff, b_nc = 0, b_fc = 0, b_nv = 0;
    for (int i = lane; i < nc; i += 32) {
      const int n = visits[fc + i];
      const int c_nc = n_children[fc + i], c_fc = first_child[fc + i];
      const double q = n > 0 ? value_sum[fc + i] / (double)n : 0.0;
      const double ucb = q + T.C * sqrt(lg / (double)(1 + n)) * (double)prior[fc + i];
      if (ucb > best) {  // strict: the first maximum wins
        best = ucb;
        best_i = i;
        b_nc = c_nc, b_fc = c_fc, b_nv = n;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, best_i, o);
      const int o_nc = __shfl_xor_sync(FULL, b_nc, o), o_fc = __shfl_xor_sync(FULL, b_fc, o), o_nv = __shfl_xor_sync(FULL, b_nv, o);
      if (oi != 0x7fffffff && (best_i == 0x7fffffff || ob > best || (ob == best && oi < best_i))) {
        best = ob;
        best_i = oi;
        b_nc = o_nc, b_fc = o_fc, b_nv = o_nv;
      }
    }
    if (best_i == 0x7fffffff) {  // "Failed to select a child." (node.cpp:72-75)
      if (lane == 0) {
        T.error[g] |= ERR_NO_CHILD;
        T.leaf[g] = -1;
        T.dropped[g] = 1;
      }
      return;
    }
    node = fc + best_i;
    nc = b_nc, fc = b_fc, nv = b_nv;
  }
  // materialise the leaf's board: parent's board + MakeMove(Move(flat_index)) (node.cpp:87-92)
  uint8_t *pool = T.boards + (size_t)g * T.board_cap * G::REC;
  int bi = T.board_idx[slab + node];
  uint8_t *b = recs[wib];
  if (bi < 0) {
    const int pb = T.board_idx[slab + T.parent[slab + node]];
    if (lane < G::REC / 16) reinterpret_cast<uint4 *>(b)[lane] = reinterpret_cast<const uint4 *>(pool + (size_t)pb * G::REC)[lane];
    __syncwarp();
    int ok = 1;
    if (lane == 0) {
      int from, to;
      decode_flat_move<G>(T.move_flat[slab + node], from, to);
      ok = apply_move_record<G>(b, from, to, NO_PIECE, G::NSQ, G::NSQ, 0) ? 1 : 0;
    }
    ok = __shfl_sync(FULL, ok, 0);
    bi = T.n_boards[g];
    if (bi >= T.board_cap || !ok) {
      if (lane == 0) {
        T.error[g] |= ok ? ERR_BOARD_CAP : ERR_MOVE;
        T.leaf[g] = -1;
        T.dropped[g] = 1;
      }
      return;
    }
    __syncwarp();
    if (lane < G::REC / 16) reinterpret_cast<uint4 *>(pool + (size_t)bi * G::REC)[lane] = reinterpret_cast<const uint4 *>(b)[lane];
    if (lane == 0) {
      T.board_idx[slab + node] = bi;
      T.n_boards[g] = bi + 1;
    }
  } else {
    if (lane < G::REC / 16) reinterpret_cast<uint4 *>(b)[lane] = reinterpret_cast<const uint4 *>(pool + (size_t)bi * G::REC)[lane];
    __syncwarp();
  }
  if (lane < G::REC / 16) reinterpret_cast<uint4 *>(T.leaf_boards + (size_t)g * G::REC)[lane] = reinterpret_cast<const uint4 *>(b)[lane];
  if (lane == 0) {
    T.leaf[g] = node;
    T.k[g] = b[G::OFF_TURN] & 3;
  }
}

// The reference rotates a whole leaf batch by the colour of states[0] (src/cpp/board.cpp:354-355,
// mcts.py:69), states = the leaves that are NOT terminal (mcts.py:18-26 drops terminal leaves before the batch is
// built): k[g] <- side to move of the first leaf whose status (already computed) says IN_PROGRESS.
template <class G>
__global__ void __launch_bounds__(1024) tree_batch_k_kernel(const fpc_tree T) {
  __shared__ int first;
  if (threadIdx.x == 0) first = 0x7fffffff;
  __syncthreads();
  int mine = 0x7fffffff;
  for (int g = threadIdx.x; g < T.n_games; g += blockDim.x)
    if (T.leaf[g] >= 0 && (T.leaf_status[g] & FPC_STATUS_RESULT_MASK) == 0) {
      mine = g;
      break;
    }
  if (mine != 0x7fffffff) atomicMin(&first, mine);
  __syncthreads();
  if (first == 0x7fffffff) return;
  const int k = T.leaf_boards[(size_t)first * G::REC + G::OFF_TURN] & 3;
  for (int g = threadIdx.x; g < T.n_games; g += blockDim.x) T.k[g] = k;
}

// mcts.py:66-79 + 82-89 and node.cpp:33-43,133-154, one CTA per game.
constexpr int EXP_THREADS = 256;  // 8 CTAs per SM: 1,024 games are one wave on 148 SMs
constexpr int EXP_WARPS = EXP_THREADS / 32;

// online soft-max statistics: running maximum m and sum s of exp(x - m)
__device__ __forceinline__ void softmax_merge(float &m, float &s, float om, float os) {
  const float nm = fmaxf(m, om);
  s = (m == nm ? s : s * __expf(m - nm)) + (om == nm ? os : os * __expf(om - nm));
  m = nm;
}

template <class G>
__global__ void __launch_bounds__(EXP_THREADS, 8) tree_expand_backup_kernel(const fpc_tree T, const float *logits, const float *values) {
  __shared__ float red_m[EXP_WARPS], red_s[EXP_WARPS];
  __shared__ double red_d[EXP_WARPS];
  __shared__ int red_i[EXP_WARPS];
  __shared__ float s_max, s_sum;
  __shared__ double s_msum;
  __shared__ int s_base;
  constexpr int ANC_MAX = 64;
  __shared__ int s_anc[ANC_MAX];
  __shared__ int s_nanc, s_anc_next;
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // every per-game scalar is requested at once (one memory latency instead of a chain of them), and so are the legal
  // moves (rows past the count hold stale moves of earlier leaves: valid memory, ignored below)
  const int leaf = T.leaf[g];
  const int status = T.leaf_status[g];
  const int cnt = T.leaf_counts[g];
  const int k = T.k[g] & 3;
  const float value_in = values[g];
  constexpr int PER = (MAX_MOVES + EXP_THREADS - 1) / EXP_THREADS;
  int flat_in[PER], prev_in[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int i = tid + q * EXP_THREADS;
    flat_in[q] = i < MAX_MOVES ? T.leaf_flat[(size_t)g * MAX_MOVES + i] : -1;
    prev_in[q] = (i > 0 && i < MAX_MOVES) ? T.leaf_flat[(size_t)g * MAX_MOVES + i - 1] : -1;
  }
  if (leaf < 0) return;
  const size_t slab = (size_t)g * T.node_cap;
  // The latency-bound tail is prepared while the logits stream: one thread (of the last warp) walks the leaf's ancestors
  // now -- a chain of dependent loads -- so that the backup at the end is one parallel round of updates; thread 0 has
  // the node count of the slab in a register by then.
  if (tid == EXP_THREADS - 32) {
    int node = leaf, cnt = 0;
    while (node >= 0 && cnt < ANC_MAX) {
      s_anc[cnt++] = node;
      node = T.parent[slab + node];
    }
    s_nanc = cnt;
    s_anc_next = node;  // >= 0 only for a path longer than ANC_MAX
  }
  const int nodes_before = tid == 0 ? T.n_nodes[g] : 0;
  const int result = status & FPC_STATUS_RESULT_MASK;
  float value;
  if (result != 0) {
    // terminal leaf: Backpropagate(0) for a stalemate, Backpropagate(-1) for any win; the root drops out
    value = result == 3 ? 0.0f : -1.0f;
    if (tid == 0) T.dropped[g] = 1;
  } else {
    value = value_in;
    const float *lg = logits + (size_t)g * G::ASZ;
    // ---- the legal moves' own logits are requested first (scattered, ~20 of them) ---------------------
    int flat[PER];
    float lgt[PER];
    bool first[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = tid + q * EXP_THREADS;
      flat[q] = -1, lgt[q] = 0.0f, first[q] = false;
      if (i < cnt) {
        flat[q] = flat_in[q];
        first[q] = i == 0 || prev_in[q] != flat[q];  // promotions share an index
        if (first[q]) {
          // policy[plane][r][c] = rot90(net_out, -k)[plane][r][c]: one clockwise quarter turn reads out[i][j] = in[R-1-j][i]
          const int plane = flat[q] / G::NSQ, sq = flat[q] - plane * G::NSQ;
          int r = sq / G::R, c = sq - r * G::R;
          for (int t = 0; t < k; ++t) {
            const int nr = G::R - 1 - c;
            c = r;
            r = nr;
          }
          lgt[q] = __ldg(lg + plane * G::NSQ + r * G::R + c);
        }
      }
    }
    // ---- softmax over the whole action space (mcts.py:67): ONE streaming pass, online max / sum; four independent
    //      16-byte loads per thread and iteration keep enough bytes in flight to stream at the HBM rate -------------
    float m = -CUDART_INF_F, s = 0.0f;
    const float4 *lg4 = reinterpret_cast<const float4 *>(lg);
    constexpr int N4 = G::ASZ / 4;
    int i = tid;
    for (; i + 3 * EXP_THREADS < N4; i += 4 * EXP_THREADS) {
      const float4 a = __ldcs(lg4 + i), b = __ldcs(lg4 + i + EXP_THREADS), c = __ldcs(lg4 + i + 2 * EXP_THREADS),
                   d = __ldcs(lg4 + i + 3 * EXP_THREADS);
      const float mx = fmaxf(fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))),
                             fmaxf(fmaxf(fmaxf(c.x, c.y), fmaxf(c.z, c.w)), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w))));
      const float nm = fmaxf(m, mx);
      s = s * __expf(m - nm) + ((__expf(a.x - nm) + __expf(a.y - nm)) + (__expf(a.z - nm) + __expf(a.w - nm))) +
          ((__expf(b.x - nm) + __expf(b.y - nm)) + (__expf(b.z - nm) + __expf(b.w - nm))) +
          ((__expf(c.x - nm) + __expf(c.y - nm)) + (__expf(c.z - nm) + __expf(c.w - nm))) +
          ((__expf(d.x - nm) + __expf(d.y - nm)) + (__expf(d.z - nm) + __expf(d.w - nm)));
      m = nm;
    }
    for (; i < N4; i += EXP_THREADS) {
      const float4 a = __ldcs(lg4 + i);
      const float nm = fmaxf(m, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
      s = s * __expf(m - nm) + ((__expf(a.x - nm) + __expf(a.y - nm)) + (__expf(a.z - nm) + __expf(a.w - nm)));
      m = nm;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) softmax_merge(m, s, __shfl_xor_sync(FULL, m, o), __shfl_xor_sync(FULL, s, o));
    if (lane == 0) {
      red_m[warp] = m;
      red_s[warp] = s;
    }
    __syncthreads();
    if (tid == 0) {
      float M = red_m[0];
      for (int w = 1; w < EXP_WARPS; ++w) M = fmaxf(M, red_m[w]);
      double S = 0.0;
      for (int w = 0; w < EXP_WARPS; ++w)
        if (red_s[w] > 0.0f) S += (double)red_s[w] * (double)expf(red_m[w] - M);
      s_max = M;
      s_sum = (float)S;
    }
    __syncthreads();
    // ---- legal moves: un-rotated above; mask, renormalise (mcts.py:69-76) ---------------------------------
    float p[PER];
    double psum = 0.0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      p[q] = first[q] ? expf(lgt[q] - s_max) / s_sum : 0.0f;
      psum += (double)p[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(FULL, psum, o);
    if (lane == 0) red_d[warp] = psum;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < EXP_WARPS; ++w) t += red_d[w];
      s_msum = t;
    }
    __syncthreads();
    // ---- one child per non-zero prior, ascending flat index (mcts.py:83-87, node.cpp:79-98): move i = tid + q*THREADS,
    //      so the children of slice q come after every child of slice q-1 -------------------------------------------
    int base_q = 0;
    bool capped = false;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const float prior = (float)((double)p[q] / s_msum);
      const bool make = first[q] && prior != 0.0f;
      const unsigned bal = __ballot_sync(FULL, make);
      if (lane == 0) red_i[warp] = __popc(bal);
      __syncthreads();
      int before = 0, tot = 0;
      for (int w = 0; w < EXP_WARPS; ++w) {
        const int c = red_i[w];
        if (w < warp) before += c;
        tot += c;
      }
      if (q == 0) {
        // the total number of children is needed up front: count the later slices too
        int all = tot;
#pragma unroll
        for (int q2 = 1; q2 < PER; ++q2) {
          const float pr2 = (float)((double)p[q2] / s_msum);
          all += __syncthreads_count(first[q2] && pr2 != 0.0f);
        }
        if (tid == 0) {
          int base = nodes_before;
          if (base + all > T.node_cap) {
            T.error[g] |= ERR_NODE_CAP;
            base = -1;
          } else {
            T.n_nodes[g] = base + all;
            T.first_child[slab + leaf] = base;
            T.n_children[slab + leaf] = all;
          }
          s_base = base;
        }
        __syncthreads();
        capped = s_base < 0;
      }
      if (make && !capped) {
        const size_t c = slab + s_base + base_q + before + __popc(bal & ((1u << lane) - 1u));
        T.parent[c] = leaf;
        T.first_child[c] = 0;
        T.n_children[c] = 0;
        T.visits[c] = 1;  // node.h:28
        T.move_flat[c] = flat[q];
        T.board_idx[c] = -1;
        T.value_sum[c] = 0.0;
        T.prior[c] = prior;
      }
      base_q += tot;
      __syncthreads();
    }
  }
  // ---- Node::Backpropagate (node.cpp:133-142): +v at the leaf, sign flips at every ancestor ------
  __syncthreads();  // the ancestor list is complete (and, when expanding, every child has been written)
  if (tid < s_nanc) {
    const int node = s_anc[tid];
    T.value_sum[slab + node] += (tid & 1) ? -(double)value : (double)value;
    T.visits[slab + node] += 1;
  }
  if (tid == 0 && s_anc_next >= 0) {  // deeper than ANC_MAX: the rest of the path one by one
    double v = (ANC_MAX & 1) ? -(double)value : (double)value;
    int node = s_anc_next;
    while (node >= 0) {
      T.value_sum[slab + node] += v;
      T.visits[slab + node] += 1;
      v = -v;
      node = T.parent[slab + node];
    }
  }
}

#define FPC_DISPATCH(R, CALL)                                              \
  switch (R) {                                                             \
    case 14: { using G = Geo<14, 3>; CALL; break; }                        \
    case 13: { using G = Geo<13, 3>; CALL; break; }                        \
    case 10: { using G = Geo<10, 2>; CALL; break; }                        \
    case 8: { using G = Geo<8, 2>; CALL; break; }                          \
    default: return fail(FPC_ERR_ARG, "unsupported board size R=" + std::to_string(R)); \
  }

static int check_tree(const fpc_tree *t, const char *who) {
  if (!t) return fail(FPC_ERR_ARG, std::string(who) + ": null tree");
  if (!fpc_supported(t->R) || t->n_games < 0 || t->node_cap < 1 || t->board_cap < 1)
    return fail(FPC_ERR_ARG, std::string(who) + ": bad tree geometry");
  if (t->n_games > 0 &&
      (!t->parent || !t->first_child || !t->n_children || !t->visits || !t->move_flat || !t->board_idx ||
       !t->value_sum || !t->prior || !t->n_nodes || !t->n_boards || !t->leaf || !t->dropped || !t->error ||
       !t->boards || !t->leaf_boards || !t->leaf_flat || !t->leaf_counts || !t->leaf_status || !t->k))
    return fail(FPC_ERR_ARG, std::string(who) + ": null tree array");
  return FPC_OK;
}

}  // namespace fpc

using namespace fpc;

extern "C" {

int fpc_tree_reset(const fpc_tree *t, const uint8_t *d_root_boards, void *stream) {
  int rc = check_tree(t, "fpc_tree_reset");
  if (rc != FPC_OK) return rc;
  if (t->n_games == 0) return FPC_OK;
  if (!d_root_boards) return fail(FPC_ERR_ARG, "fpc_tree_reset: null roots");
  cudaStream_t st = (cudaStream_t)stream;
  FPC_DISPATCH(t->R, (tree_reset_kernel<G><<<t->n_games, 256, 0, st>>>(*t, d_root_boards)));
  return cuda_check(cudaGetLastError(), "tree_reset_kernel launch");
}

int fpc_tree_select(const fpc_tree *t, int batch_rotation, float *d_planes, int flags, fpc_dense_track *track, void *stream) {
  int rc = check_tree(t, "fpc_tree_select");
  if (rc != FPC_OK) return rc;
  if (t->n_games == 0) return FPC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (t->n_games + SEL_WARPS - 1) / SEL_WARPS;
  FPC_DISPATCH(t->R, (tree_descend_kernel<G><<<blocks, SEL_WARPS * 32, 0, st>>>(*t)));
  CK(cudaGetLastError());
  if (batch_rotation) {
    // the batch rotation is that of the first NON-terminal leaf: results first, then k, then the planes
    rc = fpc_observe(t->R, t->leaf_boards, t->n_games, nullptr, t->leaf_flat, t->leaf_counts, t->leaf_status, nullptr,
                     nullptr, 0, nullptr, 0, stream);
    if (rc != FPC_OK) return rc;
    FPC_DISPATCH(t->R, (tree_batch_k_kernel<G><<<1, 1024, 0, st>>>(*t)));
    CK(cudaGetLastError());
    return fpc_observe_tracked(track, t->R, t->leaf_boards, t->n_games, nullptr, nullptr, nullptr, nullptr, d_planes, t->k, 0,
                               nullptr, flags, stream);
  }
  // legal moves + GetGameResult + planes of the leaf batch: the environment's rules kernel
  return fpc_observe_tracked(track, t->R, t->leaf_boards, t->n_games, nullptr, t->leaf_flat, t->leaf_counts, t->leaf_status,
                             d_planes, t->k, 0, nullptr, flags, stream);
}

int fpc_tree_expand_backup(const fpc_tree *t, const float *d_logits, const float *d_values, void *stream) {
  int rc = check_tree(t, "fpc_tree_expand_backup");
  if (rc != FPC_OK) return rc;
  if (t->n_games == 0) return FPC_OK;
  if (!d_logits || !d_values) return fail(FPC_ERR_ARG, "fpc_tree_expand_backup: null logits / values");
  if (reinterpret_cast<uintptr_t>(d_logits) & 15) return fail(FPC_ERR_ARG, "logits must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  FPC_DISPATCH(t->R, (tree_expand_backup_kernel<G><<<t->n_games, EXP_THREADS, 0, st>>>(*t, d_logits, d_values)));
  return cuda_check(cudaGetLastError(), "tree_expand_backup_kernel launch");
}

}  // extern "C"

// TEST INFRASTRUCTURE.  A one-warp SIMT emulator for the host: the product's warp-level kernel bodies
// (csrc/fpc_rules.cuh) are compiled with g++ and every lane runs as a ucontext fibre.  Between two warp
// collectives (__ballot_sync, __shfl_sync, __any_sync, __syncwarp ...) the lanes run one after the other, in an
// order that alternates from collective to collective, so a missing __syncwarp() between a shared-memory write
// and another lane's read shows up as a wrong result; a collective that not every live lane reaches (divergent
// call sites) aborts with a message.  Nothing here is used by the product.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>

namespace warp_emul {

constexpr int LANES = 32;
enum Kind { K_NONE = 0, K_BALLOT, K_ANY, K_ALL, K_SYNC, K_SHFL, K_SHFL_XOR, K_SHFL_UP, K_SHFL64 };

struct Warp {
  ucontext_t main_ctx, ctx[LANES];
  char *stack[LANES];
  bool done[LANES];
  int cur = -1;
  long sweeps = 0;
  // contributions of the collective every live lane is waiting in
  int kind[LANES];
  int site[LANES];
  uint64_t in_val[LANES];
  int in_arg[LANES];
  // results, valid during the following sweep
  uint32_t out_ballot = 0;
  uint64_t out_val[LANES];
  std::function<void(int)> body;
};

inline Warp &W() {
  static Warp w;
  return w;
}

inline void lane_entry() {
  Warp &w = W();
  const int lane = w.cur;
  w.body(lane);
  w.done[lane] = true;
  w.kind[lane] = K_NONE;
  swapcontext(&w.ctx[lane], &w.main_ctx);
}

inline void collective_wait(int kind, int site, uint64_t val, int arg) {
  Warp &w = W();
  const int lane = w.cur;
  w.kind[lane] = kind;
  w.site[lane] = site;
  w.in_val[lane] = val;
  w.in_arg[lane] = arg;
  swapcontext(&w.ctx[lane], &w.main_ctx);
}

// Runs body(lane) for 32 lanes to completion.
inline void run_warp(const std::function<void(int)> &body) {
  Warp &w = W();
  constexpr size_t STACK = 256 * 1024;
  w.body = body;
  w.sweeps = 0;
  for (int l = 0; l < LANES; ++l) {
    if (!w.stack[l]) w.stack[l] = (char *)malloc(STACK);
    getcontext(&w.ctx[l]);
    w.ctx[l].uc_stack.ss_sp = w.stack[l];
    w.ctx[l].uc_stack.ss_size = STACK;
    w.ctx[l].uc_link = &w.main_ctx;
    makecontext(&w.ctx[l], (void (*)())lane_entry, 0);
    w.done[l] = false;
    w.kind[l] = K_NONE;
  }
  for (;;) {
    int live = 0;
    const bool rev = (w.sweeps & 1) != 0;
    for (int i = 0; i < LANES; ++i) {
      const int l = rev ? LANES - 1 - i : i;
      if (w.done[l]) continue;
      w.cur = l;
      swapcontext(&w.main_ctx, &w.ctx[l]);
      if (!w.done[l]) ++live;
    }
    ++w.sweeps;
    if (!live) break;
    // every live lane now waits in a collective: they must all be the same one
    int k = K_NONE, st = -1;
    for (int l = 0; l < LANES; ++l) {
      if (w.done[l]) continue;
      if (k == K_NONE) k = w.kind[l], st = w.site[l];
      if (w.kind[l] != k || w.site[l] != st) {
        fprintf(stderr, "warp_emul: divergent collective: lane %d at kind %d site %d, others at kind %d site %d\n", l,
                w.kind[l], w.site[l], k, st);
        abort();
      }
    }
    uint32_t bal = 0;
    for (int l = 0; l < LANES; ++l)
      if (!w.done[l] && w.in_val[l]) bal |= 1u << l;
    uint32_t live_mask = 0;
    for (int l = 0; l < LANES; ++l)
      if (!w.done[l]) live_mask |= 1u << l;
    switch (k) {
      case K_BALLOT: w.out_ballot = bal; break;
      case K_ANY: w.out_ballot = bal != 0; break;
      case K_ALL: w.out_ballot = (bal & live_mask) == live_mask; break;
      case K_SYNC: break;
      case K_SHFL:
      case K_SHFL64:
        for (int l = 0; l < LANES; ++l) w.out_val[l] = w.in_val[w.in_arg[l] & 31];
        break;
      case K_SHFL_XOR:
        for (int l = 0; l < LANES; ++l) w.out_val[l] = w.in_val[(l ^ w.in_arg[l]) & 31];
        break;
      case K_SHFL_UP:
        for (int l = 0; l < LANES; ++l) w.out_val[l] = l - w.in_arg[l] >= 0 ? w.in_val[l - w.in_arg[l]] : w.in_val[l];
        break;
      default: fprintf(stderr, "warp_emul: bad collective %d\n", k); abort();
    }
  }
}

}  // namespace warp_emul

// ---- the CUDA spellings the kernel bodies use ------------------------------------------------------------
#define FPC_SITE __LINE__
inline unsigned __ballot_sync_(int site, unsigned, int pred) {
  warp_emul::collective_wait(warp_emul::K_BALLOT, site, pred != 0, 0);
  return warp_emul::W().out_ballot;
}
inline int __any_sync_(int site, unsigned, int pred) {
  warp_emul::collective_wait(warp_emul::K_ANY, site, pred != 0, 0);
  return (int)warp_emul::W().out_ballot;
}
inline int __all_sync_(int site, unsigned, int pred) {
  warp_emul::collective_wait(warp_emul::K_ALL, site, pred != 0, 0);
  return (int)warp_emul::W().out_ballot;
}
inline void __syncwarp_(int site) { warp_emul::collective_wait(warp_emul::K_SYNC, site, 0, 0); }
template <class T>
inline T __shfl_sync_(int site, unsigned, T v, int src) {
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  warp_emul::collective_wait(warp_emul::K_SHFL, site, raw, src);
  raw = warp_emul::W().out_val[warp_emul::W().cur];
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
template <class T>
inline T __shfl_xor_sync_(int site, unsigned, T v, int m) {
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  warp_emul::collective_wait(warp_emul::K_SHFL_XOR, site, raw, m);
  raw = warp_emul::W().out_val[warp_emul::W().cur];
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
#define __ballot_sync(m, p) __ballot_sync_(FPC_SITE, (m), (p))
#define __any_sync(m, p) __any_sync_(FPC_SITE, (m), (p))
#define __all_sync(m, p) __all_sync_(FPC_SITE, (m), (p))
#define __syncwarp() __syncwarp_(FPC_SITE)
#define __shfl_sync(m, v, s) __shfl_sync_(FPC_SITE, (m), (v), (s))
#define __shfl_xor_sync(m, v, s) __shfl_xor_sync_(FPC_SITE, (m), (v), (s))

inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
template <class T>
inline T __ldg(const T *p) { return *p; }
template <class T>
inline void __stcs(T *p, T v) { *p = v; }
inline unsigned atomicOr(unsigned *p, unsigned v) {
  unsigned o = *p;
  *p = o | v;
  return o;
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  unsigned long long o = *p;
  *p = o + v;
  return o;
}

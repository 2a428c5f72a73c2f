// TEST INFRASTRUCTURE.  Compiles the product's rules kernel body (csrc/fpc_rules.cuh: rules_warp, the code every
// warp of rules_kernel runs) for the HOST with g++ and runs it under a fibre-per-lane warp emulator
// (warp_emul.h), so that the whole warp choreography -- ballots, prefix sums, shared-memory tables, atomics -- is
// checked against the oracle in the GPU-less build container.  The -m gpu tests then run the same code on a B200.
#include "warp_emul.h"

#include <cuda_runtime.h>

#include "fpc_rules.cuh"

using namespace fpc;

template <class G>
static int run(const ObserveParams &P) {
  static RulesScratch<G> scratch;
  // stale shared memory from the previous game is part of the test: nothing may depend on it
  for (int g = 0; g < P.n; ++g)
    warp_emul::run_warp([&](int lane) { rules_warp<G>(P, scratch, g, lane); });
  return 0;
}

extern "C" int emul_rules(int R, const ObserveParams *P) {
  switch (R) {
    case 14: return run<Geo<14, 3>>(*P);
    case 13: return run<Geo<13, 3>>(*P);
    case 10: return run<Geo<10, 2>>(*P);
    case 8: return run<Geo<8, 2>>(*P);
  }
  return -1;
}
extern "C" int emul_strides(int *cell_first, int *cell_stride, int *flat_first, int *flat_stride) {
  *cell_first = CELL_FIRST, *cell_stride = CELL_STRIDE, *flat_first = FLAT_FIRST, *flat_stride = FLAT_STRIDE;
  return 0;
}

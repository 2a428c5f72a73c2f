#!/bin/bash
# tuning sweep of expand_kernel launch shape (not part of the product)
for t in 128 256 512; do for i in 1 2 4 8; do
  echo -n "threads=$t iters=$i : "
  FPC_EXPAND_THREADS=$t FPC_EXPAND_ITERS=$i python tools/probe.py 2>&1 | grep -E "planes\+mask  |async" | awk '{printf "%s ", $(NF-1)}'; echo
done; done

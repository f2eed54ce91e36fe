#!/bin/bash
# component timing of the row-streaming conv: all | no MMA | no row TMA | no epilogue stores | combos
for d in 0 1 2 4 3 5 6 7; do echo "== NERVECL_ROWS_DBG=$d"; NERVECL_ROWS_DBG=$d python scripts/bench_conv.py rows 2>&1 | tail -4; done

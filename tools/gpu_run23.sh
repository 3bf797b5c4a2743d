#!/bin/bash
# last call of the round: the rebuilt binaries of the final tree — smoke() and two quick parity tests
set -x
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python -m pytest tests -m gpu -x -q -k "bit_exact_small_ring or aggregate_check or war_on_gpu" 2>&1 | tail -2

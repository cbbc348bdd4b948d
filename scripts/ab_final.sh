#!/bin/bash
# same-box A/B of the round-2 step variants (bf16, B=256, gamma=4, V=128256): default, static rows, atomics tail, one chunk
for o in "" "static_rows=1" "tail_slots=0" "tail_slots=0,tf_balance=0,static_rows=1" "chunks=1"; do echo "== ${o:-default}"; TIME=1 BS=256 OPTS=$o timeout -s KILL 80 python scripts/small_batch.py | grep B=; done

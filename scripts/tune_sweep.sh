#!/bin/bash
# tuning sweep of the library's SPECDEC_OPTS hooks on the headline bench (prints ms/step, graph-replay ms/step)
for o in "" "chunks=1" "p1_ctas=4" "p1_ctas=2" "tf_ch=16" "tf_ch=24" "chunks=3" "chunk0_pct=40" "chunk0_pct=60"; do SPECDEC_OPTS=$o timeout 200 python bench.py --no-cpu-baseline --no-sweep --e2e-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"$o\", round(d[\"ms_per_step\"],4), round(d[\"graph_replay_ms_per_step\"],4), round(d[\"roofline\"][\"kernel_ms\"],4))"; done

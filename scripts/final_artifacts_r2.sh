#!/bin/bash
# Round-2 artifacts on one B200: tests, headline + dtype benches, reference arm, ncu launch list (incl. the full-batch
# row-kernel launch the roofline fraction rests on) + full captures of the hot kernels.
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_test_final.log 2>&1; echo pytest rc=$?; tail -2 $O/r2_test_final.log
timeout 900 python bench.py > $O/r2_bench_bf16.json 2> $O/r2_bench_bf16.err; echo rc=$?
timeout 300 python bench.py --dtype f32 --no-sweep --no-cpu-baseline > $O/r2_bench_f32.json 2> $O/r2_bench_f32.err; echo rc=$?
timeout 300 python bench.py --dtype f16 --no-sweep --no-cpu-baseline > $O/r2_bench_f16.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --mode topk50 --no-sweep --no-cpu-baseline --steps 50 > $O/r2_bench_topk50.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --mode nucleus0.9 --no-sweep --no-cpu-baseline --steps 20 > $O/r2_bench_nucleus.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_reference.json 2>/dev/null; echo rc=$?
# ncu launch list of the bench command: the K timed two-chunk steps AND the single-chunk pass whose full-batch
# row-kernel launch (grid 592, 591 MB) is the one roofline.frac is computed from
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rowfast|plan_|tail_|exact_rows|sample_partial" -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline --e2e-steps 1 > $O/ncu_l.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rowfast_tma|tail_slots" -c 2 -o /tmp/r2_prof_hot env BS=256 N=2 OPTS=chunks=1 python scripts/small_batch.py > $O/ncu_f.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rowsel_tma" -c 1 -o /tmp/r2_prof_rowsel python scripts/masked_one.py > $O/ncu_m.log 2>&1; echo rc=$?
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
ncu -i /tmp/r2_prof_hot.ncu-rep --page raw --csv --metrics $M > $O/r2_ncu_full_hot.csv 2>/dev/null
ncu -i /tmp/r2_prof_rowsel.ncu-rep --page raw --csv --metrics $M > $O/r2_ncu_full_rowsel.csv 2>/dev/null
nproc; lscpu | grep "Model name"

#!/bin/bash
# Round artifacts on one B200: tests, headline + dtype benches, reference arm, ncu launch list + full captures.
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/test_final.log 2>&1; echo pytest rc=$?; tail -2 $O/test_final.log
timeout 600 python bench.py > $O/r1_bench_bf16.json 2> $O/r1_bench_bf16.err; echo rc=$?
timeout 300 python bench.py --dtype f32 --no-sweep --no-cpu-baseline > $O/r1_bench_f32.json 2> $O/r1_bench_f32.err; echo rc=$?
timeout 300 python bench.py --dtype f16 --no-sweep --no-cpu-baseline > $O/r1_bench_f16.json 2>/dev/null; echo rc=$?
timeout 300 python bench.py --mode nucleus0.9 --no-sweep --no-cpu-baseline --steps 20 > $O/r1_bench_nucleus.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r1_bench_reference.json 2>/dev/null; echo rc=$?
timeout 600 python bench.py --impl reference --mode nucleus0.9 --steps 2 --warmup 1 > $O/r1_bench_reference_nucleus0.9.json 2>/dev/null; echo rc=$?
# ncu: launch list of the bench command (our kernels), then one full capture of each hot kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rowfast|plan_|tail_fused|exact_rows|sample_partial" -c 60 --csv --log-file $O/r1_launches.csv python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline --e2e-steps 1 > $O/ncu_l.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rowfast_tma|tail_fused" -s 6 -c 2 -o /tmp/r1_prof_hot python bench.py --steps 3 --warmup 2 --no-sweep --no-cpu-baseline --e2e-steps 1 > $O/ncu_f.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nucleus_hist|nucleus_fast" -s 6 -c 2 -o /tmp/r1_prof_nucleus python bench.py --mode nucleus0.9 --steps 3 --warmup 2 --no-sweep --no-cpu-baseline --e2e-steps 1 > $O/ncu_n.log 2>&1; echo rc=$?
# the .ncu-rep files (source imported) exceed what gpurun_out may carry back: export the raw pages here
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
ncu -i /tmp/r1_prof_hot.ncu-rep --page raw --csv --metrics $M > $O/r1_ncu_full_hot.csv 2>/dev/null
ncu -i /tmp/r1_prof_nucleus.ncu-rep --page raw --csv --metrics $M > $O/r1_ncu_full_nucleus.csv 2>/dev/null
nproc; lscpu | grep "Model name"

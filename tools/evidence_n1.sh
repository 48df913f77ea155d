#!/bin/bash
# One-GPU evidence set (run on the GPU box): full GPU test suite, the bench line, the reference arm, the ncu launch list
# of the bench command, the fmaheavy pipe counters and one --set full capture of the default 768-bit kernel.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/evidence_n1.sh'
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; tail -3 gpurun_out/gputest.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_n1_reference_arm.json 2>> gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-parity > /dev/null 2>&1
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_integer_pred_on.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:ntt768_pass -c 4 --csv --log-file gpurun_out/pipes_cta_wide.csv python tools/run_ntt768_once.py 20 4 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt768_pass -s 2 -c 2 -f -o gpurun_out/prof_cta_wide python tools/run_ntt768_once.py 20 4 > /dev/null 2>&1
ls -la gpurun_out | tail -12

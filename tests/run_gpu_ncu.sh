#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) full captures of the dominant kernels.
# Each ncu command runs only after the same command has exited 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tests/bench_kernels.py rec_one 64 300 > gpurun_out/rec_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_tc -s 2 -c 1 -o gpurun_out/rec_tc_full python tests/bench_kernels.py rec_one 64 300 > gpurun_out/ncu_rec_full.log 2>&1
echo "rec full rc=$?"
python tests/bench_kernels.py gemm_one > gpurun_out/gemm_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -o gpurun_out/gemm_full python tests/bench_kernels.py gemm_one > gpurun_out/ncu_gemm_full.log 2>&1
echo "gemm full rc=$?"
python tests/bench_kernels.py attn_one 48 > gpurun_out/attn_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:band_attn_fwd -s 3 -c 1 -o gpurun_out/attn_full python tests/bench_kernels.py attn_one 48 > gpurun_out/ncu_attn_full.log 2>&1
echo "attn full rc=$?"

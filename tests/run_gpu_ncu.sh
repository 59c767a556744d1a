#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) one full capture of the dominant kernel.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tests/bench_kernels.py rec_one 64 300 > gpurun_out/rec_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_cluster -s 2 -c 1 -o gpurun_out/rec_full python tests/bench_kernels.py rec_one 64 300 > gpurun_out/ncu_rec_full.log 2>&1
echo "rec full rc=$?"
python tests/bench_kernels.py gemm_one > gpurun_out/gemm_one_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -o gpurun_out/gemm_full python tests/bench_kernels.py gemm_one > gpurun_out/ncu_gemm_full.log 2>&1
echo "gemm full rc=$?"

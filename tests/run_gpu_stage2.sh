#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "tf32x3 M=|passed|failed|FAILED" gpurun_out/tests.log | tail -30
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== rec microbench"; timeout 600 python tests/bench_kernels.py rec > gpurun_out/rec.log 2>&1; cat gpurun_out/rec.log
echo "== gemm microbench"; timeout 600 python tests/bench_kernels.py gemm > gpurun_out/gemm.log 2>&1; cat gpurun_out/gemm.log
echo "== ncu rec kernel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_fwd_cluster -s 2 -c 1 -o gpurun_out/rec_prof python tests/bench_kernels.py rec_one 64 100 > gpurun_out/ncu_rec.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_rec.log

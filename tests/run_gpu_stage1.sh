#!/bin/bash
# Staged first-contact run on the GPU box: safest kernels first, each stage under its own timeout so that a
# hung kernel cannot take the whole call (and the box) down.
mkdir -p gpurun_out
P="python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== stage A: simt gemm, crf, losses";  timeout 300 $P -k "gemm_f32 or gemm_tn or crf_golden or crf_viterbi or crf_nll or losses" > gpurun_out/stageA.log 2>&1; echo "A rc=$?"; tail -5 gpurun_out/stageA.log
echo "== stage B: golden LSTM family with SIMT projections"; MTS_GEMM_IMPL=simt timeout 300 $P -k "bilstm_golden or late_fusion or bilstm_crf or pair_input" > gpurun_out/stageB.log 2>&1; echo "B rc=$?"; tail -5 gpurun_out/stageB.log
echo "== stage C: cluster recurrence (H=256) with SIMT projections"; MTS_GEMM_IMPL=simt timeout 600 $P -k "h256 or padding_invariance or text_segmenter" > gpurun_out/stageC.log 2>&1; echo "C rc=$?"; tail -8 gpurun_out/stageC.log
echo "== stage D: tcgen05 gemm"; timeout 300 $P -k "gemm_tf32x3" > gpurun_out/stageD.log 2>&1; echo "D rc=$?"; tail -8 gpurun_out/stageD.log
echo "== stage E: everything, default path"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/stageE.log 2>&1; echo "E rc=$?"; tail -8 gpurun_out/stageE.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err

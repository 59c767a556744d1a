#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|FAILED" gpurun_out/tests.log | tail -30
echo "== rec microbench"; timeout 600 python tests/bench_kernels.py rec > gpurun_out/rec.log 2>&1; cat gpurun_out/rec.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['kernel_ms_per_call'], d['roofline']['frac'])"

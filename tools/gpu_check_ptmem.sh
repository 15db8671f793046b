#!/bin/bash
# One GPU box visit: P-in-TMEM stress test, the GPU suite under both attention variants, benches A/B.
mkdir -p gpurun_out
{
echo "--- stress"; timeout 120 python -m pytest tests/test_gpu_kernels.py -q -x -k p_in_tmem_stress 2>&1 | tail -3
echo "--- suite PTMEM=1"; ACE_ATTN_PTMEM=1 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "--- suite PTMEM=0"; ACE_ATTN_PTMEM=0 timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
} > gpurun_out/ptmem_check.log 2>&1
ACE_ATTN_PTMEM=1 timeout 200 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2_pt1.json 2>&1
ACE_ATTN_PTMEM=0 timeout 200 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2_pt0.json 2>&1
ACE_ATTN_PTMEM=1 timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_pt1.json 2>&1
ACE_ATTN_PTMEM=0 timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_pt0.json 2>&1
cat gpurun_out/ptmem_check.log
for f in gpurun_out/bench_c?_pt?.json; do echo $f; tail -c 600 $f | head -c 300; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d["value"], d["e2e"]["value"], d["config"].get("outputs_finite"))
except Exception as e: print("ERR", e)
PY
done

#!/bin/bash
# One GPU box visit at the end of a change: GPU suite, smoke, benches (default + attention A/B), ncu launch list and
# --set full captures of the attention kernel.  Everything lands in gpurun_out/.
V=${1:-v5}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/gputests.log 2>&1; tail -2 gpurun_out/gputests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
ACE_ATTN_PTMEM=0 timeout 200 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2_pt0.json 2>&1
ACE_ATTN_PTMEM=0 timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_pt0.json 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r1_${V}_launches.csv python tools/profile_step.py > gpurun_out/prof.log 2>&1
PROF_VAE=0 timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:attention_tc_kernel --launch-count 4 -f -o gpurun_out/r1_${V}_attn python tools/profile_step.py >> gpurun_out/prof.log 2>&1
PROF_VAE=0 PROF_T=6000 timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:attention_tc_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r1_${V}_attn_c3 python tools/profile_step.py >> gpurun_out/prof.log 2>&1
for f in gpurun_out/bench_c2.json gpurun_out/bench_c3.json gpurun_out/bench_c2_pt0.json gpurun_out/bench_c3_pt0.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "finite", d["config"].get("outputs_finite"), "attn_ms", d["breakdown_ms_per_song"]["attention"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done

#!/bin/bash
# One GPU box visit at the end of a change: GPU suite, smoke, benches of every single-GPU workload, the output-path
# timings, the ncu launch list of one step and --set full captures.  Everything lands in gpurun_out/.
#   bash tools/gpu_round_check.sh <tag>        AB=1: also bench c2/c3 with P through shared memory (attention A/B)
#                                              ATTN=1: also capture the attention kernel with --set full
V=${1:-v5}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/gputests.log 2>&1; tail -2 gpurun_out/gputests.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --workload c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 200 python bench.py --workload c1 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
if [ -n "$AB" ]; then
ACE_B200_LIB=$PWD/ace-step-1.5-for-windows_b200/libacestep_b200_probe.so ACE_ATTN_PTMEM=0 timeout 200 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/bench_c2_pt0.json 2>&1
ACE_B200_LIB=$PWD/ace-step-1.5-for-windows_b200/libacestep_b200_probe.so ACE_ATTN_PTMEM=0 timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/bench_c3_pt0.json 2>&1
fi
timeout 120 python tools/profile_output.py > gpurun_out/output_path_timing.log 2>&1; cat gpurun_out/output_path_timing.log
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_${V}_launches.csv python tools/profile_step.py > gpurun_out/prof.log 2>&1
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:"abs_peak_kernel|peak_scale_kernel|latent_guard_kernel|cross_probs_kernel" -f -o gpurun_out/r2_${V}_output \
  python tools/profile_output.py >> gpurun_out/prof.log 2>&1
if [ -n "$ATTN" ]; then
PROF_VAE=0 timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:attention_tc_kernel --launch-count 4 -f -o gpurun_out/r2_${V}_attn python tools/profile_step.py >> gpurun_out/prof.log 2>&1
PROF_VAE=0 PROF_T=6000 timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on \
  --kernel-name regex:attention_tc_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_${V}_attn_c3 python tools/profile_step.py >> gpurun_out/prof.log 2>&1
fi
for f in gpurun_out/bench_c2.json gpurun_out/bench_c3.json gpurun_out/bench_c1.json gpurun_out/bench_c5.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "finite", d["config"].get("outputs_finite"), "attn_ms", d["breakdown_ms_per_song"]["attention"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done

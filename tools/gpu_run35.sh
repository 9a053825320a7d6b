#!/bin/bash
# 1 GPU: per-round phase breakdown of the insert kernel, launch list and --set full capture of the bench command
mkdir -p gpurun_out
for w in rs hc14; do
for r in 0 1 2 3 4; do
KMX_PHASE_ROUND=$r timeout 600 python bench.py --workload $w --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/bench_round.log 2> gpurun_out/bench_round.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_round.log') if x.startswith('{')]
j=json.loads(l[-1]); print('$w round $r insert %.2f'%j['stage_ms']['ms_insert'], j['build_stats']['insert_phase_cycles'], j['build_stats']['batches'])
PY
done
done 2>&1 | tee gpurun_out/round_phases.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 600 $CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'insert_kernel|query_fast_kernel|query_slow_kernel|encode_kernel|count_kernel' -s 5 -c 8 -o gpurun_out/prof_r1_v3 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log

#!/bin/bash
# round 2, 1-GPU measurement run: suite, headline bench, query L2-policy A/B, RS phase diagnostics, ncu launch list + full captures
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2d_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench.err; echo "bench exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2d_bench.log') if x.startswith('{')]
j=json.loads(l[-1]); print('hc14 value %.4g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.1f ms'%j['e2e']['ms_per_step'], 'q %.4g'%j['query']['value'], j['parity']['all_ranks'], j['roofline']['frac_of_random_sector_peak'])
r=j['extra']['rs']; print('rs value %.4g'%r['value'], 'ms/step %.3f'%r['ms_per_step'], r['stage_ms'], 'e2e %.2f ms'%r['e2e']['ms_per_step'], 'q %.4g'%r['query']['value'], r['parity']['all_ranks'], r['roofline']['frac_of_random_sector_peak'])
PY
for pol in 0 1 2 3 5 7; do echo "KMX_QUERY_L2=$pol"; KMX_QUERY_L2=$pol timeout 300 python tools/query_only.py hc14 4 2>&1 | tail -2; done > gpurun_out/r2d_query_l2_ab.log 2>&1; cat gpurun_out/r2d_query_l2_ab.log
for pol in 0 7; do echo "rs KMX_QUERY_L2=$pol"; KMX_QUERY_L2=$pol timeout 300 python tools/query_only.py rs 4 2>&1 | tail -2; done >> gpurun_out/r2d_query_l2_ab.log 2>&1; tail -6 gpurun_out/r2d_query_l2_ab.log
for r in 0 1 4; do
KMX_PHASE_ROUND=$r timeout 300 python bench.py --workload rs --no-cpu-baseline --no-extra --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
j=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('rs round $r insert %.3f'%j['stage_ms']['ms_insert'], j['build_stats']['insert_phase_cycles'], j['build_stats']['batches'])"
done > gpurun_out/r2d_rs_phase_by_round.log 2>&1; cat gpurun_out/r2d_rs_phase_by_round.log
python tools/random_sector_peaks.py > gpurun_out/r2d_random_sector_peaks.log 2>&1; tail -12 gpurun_out/r2d_random_sector_peaks.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2d_launches_hc14.csv $CMD > gpurun_out/r2d_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'insert_kernel|count_kernel|encode_kernel' -s 3 -c 4 -o gpurun_out/r2d_prof_hc14 -f $CMD > gpurun_out/r2d_ncu_full.log 2>&1; echo "ncu full exit $?"; tail -3 gpurun_out/r2d_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'query_fast' -s 1 -c 1 -o gpurun_out/r2d_prof_query_hc14 -f python tools/query_only.py hc14 2 > gpurun_out/r2d_ncu_query.log 2>&1; echo "ncu query exit $?"
ls -la gpurun_out/*.ncu-rep

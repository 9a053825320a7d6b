#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for sc in 1 0; do
KMX_STREAM_CELLS=$sc timeout 600 python bench.py --workload hc14 --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/bench_hc14_sc$sc.log 2> gpurun_out/bench_hc14_sc$sc.err; echo "bench hc14 stream_cells=$sc exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_hc14_sc$sc.log') if x.startswith('{')]
j=json.loads(l[-1]); print('hc14 sc=$sc value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], j['e2e_wall_ms_steps'], 'q %.3g'%(j['query']['value']))
PY
done
timeout 600 python bench.py --workload rs --no-cpu-baseline > gpurun_out/bench_rs.log 2> gpurun_out/bench_rs.err; echo "bench rs exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_rs.log') if x.startswith('{')]
j=json.loads(l[-1]); print('rs value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], j['e2e_wall_ms_steps'], 'q %.3g'%(j['query']['value']))
PY
timeout 300 python tools/trace_e2e.py rs > gpurun_out/trace_e2e.log 2>&1; grep -A12 "READERS=4" gpurun_out/trace_e2e.log | tail -12; tail -12 gpurun_out/trace_e2e.log

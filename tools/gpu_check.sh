#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for w in rs hc14; do
timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$w.log') if x.startswith('{')]
j=json.loads(l[-1]); print('$w value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], j['build_stats']['insert_iterations'], 'e2e %.3g'%j['e2e']['value'], j['e2e_wall_ms_steps'], 'q %.3g'%(j['query']['value']), j['roofline']['frac_of_random_sector_peak'])
PY
done
timeout 300 python tools/trace_e2e.py rs > gpurun_out/trace_e2e.log 2>&1; cat gpurun_out/trace_e2e.log | tail -60

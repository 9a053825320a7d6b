#!/bin/bash
# 1 GPU: full GPU suite, locality microbenchmark, headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/random_sector_locality.py > gpurun_out/locality.log 2>&1; cat gpurun_out/locality.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')]
j=json.loads(l[-1]); print('value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], j['stage_ms'], j['build_stats']['insert_phase_cycles'], 'e2e %.3g'%j['e2e']['value'], 'q %.3g'%(j['query']['value']), j['roofline']['frac_of_random_sector_peak'])
PY

#!/bin/bash
mkdir -p gpurun_out
for sc in 3 11; do
KMX_STREAM_CELLS=$sc timeout 600 python bench.py --workload hc14 --no-cpu-baseline --steps 4 --warmup 2 > gpurun_out/bench_hc14_sc$sc.log 2> gpurun_out/bench_hc14_sc$sc.err; echo "bench hc14 stream_cells=$sc exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_hc14_sc$sc.log') if x.startswith('{')]
j=json.loads(l[-1]); print('hc14 sc=$sc value %.3g'%j['value'], 'ms/step %.2f'%j['ms_per_step'], 'insert %.2f'%j['stage_ms']['ms_insert'], j['build_stats']['insert_phase_cycles'])
PY
done

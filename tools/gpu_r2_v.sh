#!/bin/bash
# round 2: claim-bitmap size on the HC14 shape (default cap 2^25 bits per array and wanted value; a full round asks for 2^27)
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
for L in ${CLAIMS:-26 27 25}; do
KMX_CLAIM_LOG2=$L timeout 100 python bench.py --no-cpu-baseline --no-extra --steps 2 --warmup 1 > gpurun_out/r2v_bench_hc14_claim$L.log 2> gpurun_out/r2v_bench_hc14_claim$L.err; echo "claim $L exit $?"
python - <<PY
import json
l=[x for x in open('gpurun_out/r2v_bench_hc14_claim$L.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('claim $L ms/step %.2f'%j['ms_per_step'], j['wall_ms_steps'], j['stage_ms']['ms_insert'], j['build_stats']['insert_iterations'], j['build_stats']['insert_phase_cycles'][:8], {k: j['parity'].get(k) for k in ('header','km.bin','rest.bin','kmer_to_occ')})
PY
done

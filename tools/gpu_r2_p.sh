#!/bin/bash
# round 2: survivor list inside the item stream (no growing list, no host round trips between launches): GPU suite, HC14 bench,
# NA12878 shape on one GPU with the host-side timeline
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/r2p_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2p_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2p_bench_hc14_n1.log 2> gpurun_out/r2p_bench_hc14_n1.err; echo "hc14 exit $?"
KMX_TRACE=1 timeout 1000 python bench.py --workload na12878 --steps 3 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2p_bench_na12878_n1.log 2> gpurun_out/r2p_bench_na12878_n1.err; echo "na12878 n1 exit $?"
python - <<PY
import json
for w in ('hc14', 'na12878'):
    l=[x for x in open('gpurun_out/r2p_bench_%s_n1.log' % w) if x.startswith('{')]
    if l:
        j=json.loads(l[-1]); print(w, 'value %.3g'%j['value'], 'ms/step %.1f'%j['ms_per_step'], j['wall_ms_steps'], j['stage_ms'], 'e2e', j['e2e']['wall_ms_steps'], {k: j['parity'].get(k) for k in ('header','km.bin','rest.bin','kmer_to_occ')}, 'rs', (j.get('extra') or {}).get('rs', {}).get('ms_per_step'))
    else: print(open('gpurun_out/r2p_bench_%s_n1.err' % w).read()[-2500:])
PY
grep "kmx\]" gpurun_out/r2p_bench_na12878_n1.err | grep -v "took 0.00" | tail -40

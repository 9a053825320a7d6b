#!/bin/bash
# round 2: encode_kernel at 4 blocks/SM (64 registers) against 3 (72 registers)
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
for lib in kmcex_b200/libkmx.so kmcex_b200/libkmx_enc3.so; do
for w in hc14 rs; do
KMX_LIB_PATH=$PWD/$lib timeout 600 python bench.py --workload $w --no-cpu-baseline --no-extra --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
j=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('$lib $w: ms/step %.3f'%j['ms_per_step'], j['stage_ms'], 'e2e %.2f'%j['e2e']['ms_per_step'], 'parity', j['parity']['all_ranks'])"
done; done 2>&1 | tee gpurun_out/r2i_encode_blocks_ab.log

#!/bin/bash
# round 2: read-only (.nc) vs ordinary load path for the query probes of a model beyond the L2: rate and DRAM bytes per probe
mkdir -p gpurun_out
export KMX_BENCH_CACHE=/tmp/kmx_bench
for pol in 7 15 0 8; do echo "KMX_QUERY_L2=$pol"; KMX_QUERY_L2=$pol timeout 300 python tools/query_only.py hc14 4 2>&1 | tail -2; done 2>&1 | tee gpurun_out/r2f_query_ld_ab.log
for pol in 7 15; do
KMX_QUERY_L2=$pol timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sector_hit_rate.pct,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'query_fast' -s 1 -c 1 --csv --log-file gpurun_out/r2f_ncu_query_pol$pol.csv python tools/query_only.py hc14 2 > /dev/null 2>&1
grep -E "query_fast" gpurun_out/r2f_ncu_query_pol$pol.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done 2>&1 | tee -a gpurun_out/r2f_query_ld_ab.log

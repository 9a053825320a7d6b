#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/query_only.py rs 3 && \
ncu --set full --clock-control none --import-source on -k regex:'query_' -s 2 -c 2 -o gpurun_out/prof_r1_query_v3 -f python tools/query_only.py rs 2 > gpurun_out/ncu_query.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_query.log

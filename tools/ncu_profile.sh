#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one eager outer step with its device time, (2) --set full capture of
# GEMM launches of the support decoder forward.  Run under gpurun on ONE GPU; numbers under ncu are never bench values.
tag=${1:-r01v5}
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --profile-step > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mtts_gemm -s 48 -c 36 \
    -f -o gpurun_out/${tag}_gemm python bench.py --profile-step > gpurun_out/${tag}_gemm.log 2>&1
ncu -i gpurun_out/${tag}_gemm.ncu-rep --page raw --csv > gpurun_out/${tag}_gemm_raw.csv 2> /dev/null
ncu -i gpurun_out/${tag}_gemm.ncu-rep --page details --csv > gpurun_out/${tag}_gemm_details.csv 2> /dev/null
rm -f gpurun_out/${tag}_gemm.ncu-rep     # gpurun_out is capped at 64 MiB; the CSV exports are what profiles/ keeps
ls -la gpurun_out/${tag}_*

#!/bin/bash
# ncu --set full capture (with source) of selected kernels of a 256-frame C2 pass.  usage: r2_ncu.sh <tag> <kernel regex> [count]
tag=$1; rx=$2; cnt=${3:-1}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 4 -c $cnt -f -o gpurun_out/${tag} \
    python bench.py --frames 256 --unique 64 --chunk 256 --steps 1 --warmup 3 --no-cpu --no-match --no-configs > gpurun_out/ncu_${tag}.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep

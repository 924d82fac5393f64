#!/bin/bash
# ncu --set full of every kernel of ONE device pass at the bench's pass size (C2, 1024 frames).  usage: r2_ncu_pass.sh <tag> [frames]
tag=$1; fr=${2:-1024}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_level0|k_resize|k_border|k_fast|k_octree|k_blur|k_describe" -s 39 -c 13 -f -o gpurun_out/${tag} \
    python bench.py --frames $fr --unique 256 --chunk $fr --steps 1 --warmup 3 --no-cpu --no-match --no-configs > gpurun_out/ncu_${tag}.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep

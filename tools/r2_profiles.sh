#!/bin/bash
# Round-2 profile set (run only after tests + bench have exited 0 without ncu):
#   <tag>_launches.csv   ncu launch list (gpu__time_duration.sum) of a short default-shaped bench run (C2 passes + e2e chunks + matching)
#   <tag>_c2_pass        --set full of every kernel of ONE 1024-frame C2 device pass (the pass size bench.py times)
#   <tag>_c5_pass        --set full of every kernel of ONE 16-frame C5 (3840x2160 / 8000 / 12) device pass
#   <tag>_matching       --set full of the all-pairs / top-2 kernels
# usage: gpurun --timeout 1500 -- bash tools/r2_profiles.sh TAG [part: a = launch list + C2 pass, b = C5 pass + matching, ab = both]
# (gpurun brings back at most 64 MiB per call: the four reports together are close to that, so the parts can be run as two calls.)
# A C2 pass is 12 launches (level 0, 7 resizes, FAST, octree, blur, describe); a C5 pass 16, and 8 launches precede the first C5 pass.
tag=${1:-r2}; part=${2:-ab}; n2=12; n5=16
mkdir -p gpurun_out
if [[ $part == *a* ]]; then
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/ncu_${tag}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_level0|k_resize|k_fast|k_octree|k_blur|k_describe" -s $((3 * n2)) -c $n2 -f -o gpurun_out/${tag}_c2_pass \
    python bench.py --frames 1024 --unique 256 --chunk 1024 --steps 1 --warmup 3 --no-cpu --no-match --no-configs > gpurun_out/ncu_${tag}_c2.log 2>&1
fi
if [[ $part == *b* ]]; then
timeout 600 ncu --set full --clock-control none --import-source on -s $((3 * n5 + 8)) -c $n5 -f -o gpurun_out/${tag}_c5_pass \
    python bench.py --only-config C5 --c5-frames 16 --c5-batch 16 --steps 1 --no-cpu > gpurun_out/ncu_${tag}_c5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_allpairs|k_top2" -c 3 -f -o gpurun_out/${tag}_matching \
    python bench.py --frames 256 --unique 64 --chunk 256 --steps 1 --warmup 1 --no-cpu --no-configs --match-q 8 --match-db 64 > gpurun_out/ncu_${tag}_matching.log 2>&1
fi
ls -la gpurun_out/${tag}_*

#!/bin/bash
# Turn the reports tools/r2_profiles.sh brought back (gpurun_out/<tag>_*.ncu-rep, <tag>_launches.csv) into the tracked files of profiles/.
# usage: bash tools/r2_profiles_emit.sh TAG
tag=${1:-r2g}; g=gpurun_out
C2='ncu --set full --clock-control none --import-source on -k regex:"k_level0|k_resize|k_fast|k_octree|k_blur|k_describe" -s 36 -c 12 python bench.py --frames 1024 --unique 256 --chunk 1024 --steps 1 --warmup 3 --no-cpu --no-match --no-configs'
C5='ncu --set full --clock-control none --import-source on -s 56 -c 16 python bench.py --only-config C5 --c5-frames 16 --c5-batch 16 --steps 1 --no-cpu'
MA='ncu --set full --clock-control none --import-source on -k regex:"k_allpairs|k_top2" -c 3 python bench.py --frames 256 --unique 64 --chunk 256 --steps 1 --warmup 1 --no-cpu --no-configs --match-q 8 --match-db 64'
rm -f profiles/ncu_constants.json
python tools/ncu_summary.py --json profiles/ncu_constants.json $g/${tag}_c2_pass.ncu-rep c2_pass 1024 "$C2"
python tools/ncu_summary.py --json profiles/ncu_constants.json $g/${tag}_c5_pass.ncu-rep c5_pass 16 "$C5"
python tools/ncu_summary.py --json profiles/ncu_constants.json $g/${tag}_matching.ncu-rep matching 2048000000 "$MA   (units = 8 x 64 keyframe pairs x 2000 x 2000 descriptor pairs per launch)"
{ echo "# C2 (640x480 / 1000 / 8 levels): every kernel of ONE 1024-frame device pass, final code of round 2"; echo; echo "Command (tools/r2_profiles.sh, part a): \`$C2\`"; echo "(the pass size bench.py times; 12 launches per pass; constants of this capture: profiles/ncu_constants.json, label c2_pass)"; echo; python tools/ncu_summary.py $g/${tag}_c2_pass.ncu-rep; } > profiles/r2_c2_pass_ncu_full.md
{ echo "# C5 (3840x2160 / 8000 features / 12 levels): every kernel of ONE 16-frame device pass, final code of round 2"; echo; echo "Command (tools/r2_profiles.sh, part b): \`$C5\`"; echo "(16 launches per pass: level 0, 11 resizes, FAST, blur forked next to the octree, describe; label c5_pass in profiles/ncu_constants.json)"; echo; python tools/ncu_summary.py $g/${tag}_c5_pass.ncu-rep; } > profiles/r2_c5_pass_ncu_full.md
{ echo "# All-pairs Hamming matching kernel, final code of round 2"; echo; echo "Command (tools/r2_profiles.sh, part b): \`$MA\`"; echo; python tools/ncu_summary.py $g/${tag}_matching.ncu-rep; } > profiles/r2_matching_ncu_full.md
python tools/ncu_srclines.py $g/${tag}_c2_pass.ncu-rep k_fast_tma 0.6 > profiles/r2_k_fast_tma_source_lines.txt
python tools/ncu_srclines.py $g/${tag}_c2_pass.ncu-rep k_describe 0.8 > profiles/r2_k_describe_source_lines.txt
python tools/ncu_srclines.py $g/${tag}_c2_pass.ncu-rep k_blur 1.0 > profiles/r2_k_blur_source_lines.txt
python - "$g/${tag}_launches.csv" <<'PY'
import csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
h = rows[0]; i = h.index('Kernel Name'); v = h.index('Metric Value'); g = h.index('Grid Size'); b = h.index('Block Size'); u = h.index('Metric Unit')
tot = {}
with open('profiles/r2_launch_list.csv', 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs  (tools/r2_profiles.sh part a; columns reduced)\n')
    f.write('id,kernel,grid,block,duration_us\n')
    for k, r in enumerate(rows[1:]):
        if len(r) <= v:
            continue
        d = float(r[v].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(r[u], 1e-3)
        name = r[i].split('(')[0].replace('void ', '')
        f.write(f'{k},{name},"{r[g]}","{r[b]}",{d:.3f}\n')
        tot[name] = tot.get(name, 0) + d
s = sum(tot.values())
for k, t in sorted(tot.items(), key=lambda x: -x[1])[:7]:
    print(f'{k:28s} {t / 1e3:9.2f} ms  {100 * t / s:5.1f} %')
PY
bash tools/sass_evidence.sh > profiles/r2_sass_evidence.txt 2>&1

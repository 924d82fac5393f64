#!/bin/bash
# Counts, per kernel of liborb_b200.so, the SASS mnemonics that prove the design claims of DESIGN.md: TMA tile loads (UTMALDG) and
# mbarrier waits (SYNCS) in k_fast_tma, integer dot products (IDP) in blur / resize, packed 3-input min/max (VIMNMX3) in FAST,
# carry-save LOP3 (LUT 0x96 / 0xe8) and POPC in the matchers.   usage: tools/sass_evidence.sh > profiles/<name>.txt
cuobjdump -sass "$(dirname "$0")/../orbslam_mapsave_b200/liborb_b200.so" 2>/dev/null |
  awk '/Function :/{fn=$3} /UTMALDG/{n["UTMALDG " fn]++} /SYNCS/{n["SYNCS " fn]++} / POPC /{n["POPC " fn]++} / IDP/{n["IDP " fn]++}
       /VIMNMX3/{n["VIMNMX3 " fn]++} /LOP3.LUT.*0x96|LOP3.LUT.*0xe8/{n["LOP3(csa) " fn]++} END{for (k in n) print n[k], k}' |
  sort -k2,2 -k1,1nr | c++filt | cut -c1-160

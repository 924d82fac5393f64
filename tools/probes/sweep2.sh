for lib in orbslam_mapsave_b200/liborb_b200.so tools/probes/liborb_fw8.so; do for p in 1 2 3 4 5 6 7; do
 ORB_B200_LIB=$lib ORBX_FW_PERSM=$p python bench.py --no-cpu --no-match --steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib persm=$p', round(d['value']), round(d['roofline']['stage_ms_per_step']['fast'],2))"
done; done

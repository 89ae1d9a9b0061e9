for d in 0 128 0 128; do B200DOCK_DBG=$d B200DOCK_LIB=diffbindfr_b200/libb200dock_trace.so timeout 200 python bench.py --no-cpu-baseline --no-sustained --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('dbg',$d, d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"; done

for i in 1 2; do timeout 200 python bench.py --no-cpu-baseline --no-sustained --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['parity']['final_lig_rmsd_A'], d['clocks']['sm_mhz'])"; done
timeout 100 python tools/timeline.py 11 2>&1 | tail -8

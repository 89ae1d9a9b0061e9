timeout 200 python bench.py --no-cpu-baseline --no-sustained --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['ms_per_step']-d['roofline']['kernel_ms_per_step'], d['value'], d['parity']['final_lig_rmsd_A'])"

timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for nu in 0 0; do B200DOCK_SKIP_NU=$nu timeout 200 python bench.py --no-cpu-baseline --no-sustained --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('skip',$nu, d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['ms_per_step']-d['roofline']['kernel_ms_per_step'], d['value'], d['parity']['final_lig_rmsd_A'])"; done

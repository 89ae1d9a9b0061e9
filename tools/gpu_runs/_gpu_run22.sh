for nu in 0 3 0 3; do B200DOCK_SKIP_NU=$nu timeout 200 python bench.py --no-cpu-baseline --no-sustained --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('skip',$nu, d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['ms_per_step']-d['roofline']['kernel_ms_per_step'], d['value'])"; done

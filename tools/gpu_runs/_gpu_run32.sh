timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --no-cpu-baseline --no-sustained --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['parity']['final_lig_rmsd_A'], d['roofline']['traffic'])"

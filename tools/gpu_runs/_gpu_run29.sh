export B200DOCK_TEST_KERNELS=11
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or golden_trajectory or bench_batch" 2>&1 | tail -2
for i in 1 2; do timeout 200 python bench.py --no-cpu-baseline --no-sustained --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['parity']['final_lig_rmsd_A'], d['clocks']['sm_mhz'])"; done

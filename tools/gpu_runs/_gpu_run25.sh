timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b25.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/b25.json'));print(d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['ms_per_step']-d['roofline']['kernel_ms_per_step'], d['value'], d['e2e']['value'], d['parity']['final_lig_rmsd_A'], d['roofline']['sustained']['steps_per_s'])"

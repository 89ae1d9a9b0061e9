timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b21.json 2> gpurun_out/b21.err || tail -5 gpurun_out/b21.err
python -c "
import json;d=json.load(open('gpurun_out/b21.json'));print(d['config']['conv_kernel'], d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['e2e']['value'], d['parity'], d['roofline'].get('sustained',{}).get('steps_per_s'))"

timeout 300 python -m pytest tests/test_noise.py tests/test_pose_init.py tests/test_mdn.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --workload cfg3_16x40 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/g_cfg3.json 2> gpurun_out/g_cfg3.err || tail -8 gpurun_out/g_cfg3.err; python -c "
import json;d=json.load(open('gpurun_out/g_cfg3.json'));print('cfg3', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('mdn'))"

for w in 3dbs 3dbs_x40; do timeout 300 python bench.py --workload $w --steps 20 --warmup 2 --no-sustained > gpurun_out/w_$w.json 2> gpurun_out/w_$w.err || tail -5 gpurun_out/w_$w.err; python -c "
import json;d=json.load(open('gpurun_out/w_$w.json'));print('$w', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity'], d['cpu_baseline']['value'])"; done
timeout 400 python bench.py --workload cfg3_16x40 --steps 20 --warmup 2 > gpurun_out/w_cfg3.json 2> gpurun_out/w_cfg3.err || tail -8 gpurun_out/w_cfg3.err; python -c "
import json;d=json.load(open('gpurun_out/w_cfg3.json'));print('cfg3', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity'], d.get('mdn'), d['config']['rank_device_ms'])"
timeout 400 python bench.py --workload posebusters_256x40 --complexes 16 --steps 20 --warmup 2 --no-cpu-baseline > gpurun_out/w_pb16.json 2> gpurun_out/w_pb16.err || tail -8 gpurun_out/w_pb16.err; python -c "
import json;d=json.load(open('gpurun_out/w_pb16.json'));print('pb16', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('mdn'))"
timeout 400 python bench.py --workload revdock_512x40 --complexes 16 --steps 20 --warmup 2 --no-cpu-baseline > gpurun_out/w_rd16.json 2> gpurun_out/w_rd16.err || tail -8 gpurun_out/w_rd16.err; python -c "
import json;d=json.load(open('gpurun_out/w_rd16.json'));print('rd16', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('mdn'))"

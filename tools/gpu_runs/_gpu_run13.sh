# r02 evidence run: wait trace (mode 6), full ncu of the conv launches, launch list, posebusters / revdock single-GPU shards
timeout 200 python tools/trace_waits.py 4 6 > gpurun_out/trace6.log 2>&1; tail -20 gpurun_out/trace6.log
timeout 400 ncu --set full --clock-control none -k regex:k_conv_fused16x2 -c 7 -o gpurun_out/prof_r02_fused16x2 -f python bench.py --no-cpu-baseline --no-sustained --steps 1 --warmup 1 > gpurun_out/ncu13.log 2>&1; tail -1 gpurun_out/ncu13.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_mode6.csv python bench.py --no-cpu-baseline --no-sustained --steps 2 --warmup 1 > gpurun_out/ncu13b.log 2>&1; tail -1 gpurun_out/ncu13b.log | cut -c1-200
timeout 500 python bench.py --workload posebusters_256x40 --complexes 16 --steps 20 --warmup 2 --no-cpu-baseline > gpurun_out/w_pb16.json 2> gpurun_out/w_pb16.err || tail -8 gpurun_out/w_pb16.err; python -c "
import json;d=json.load(open('gpurun_out/w_pb16.json'));print('pb16', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('mdn'))"

# full GPU suite with kernel 11 as the default, then the r02 evidence for it
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/b20_default.json 2> gpurun_out/b20_default.err || tail -5 gpurun_out/b20_default.err
python -c "
import json;d=json.load(open('gpurun_out/b20_default.json'));print(d['config']['conv_kernel'], d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['final_lig_rmsd_A'], d['roofline'].get('sustained',{}).get('steps_per_s'), d['roofline']['frac'], d['cpu_baseline'])"
timeout 400 ncu --set full --clock-control none -k regex:k_conv_fused16x2 -c 7 -o gpurun_out/prof_r02_split -f python bench.py --no-cpu-baseline --no-sustained --steps 1 --warmup 1 > gpurun_out/ncu20.log 2>&1; tail -1 gpurun_out/ncu20.log | cut -c1-120
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r02_k11.csv python bench.py --no-cpu-baseline --no-sustained --steps 2 --warmup 1 > gpurun_out/ncu20b.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02_cfg3.csv python bench.py --workload cfg3_16x40 --complexes 2 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu20c.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/b20_reference.json 2> gpurun_out/b20_reference.err; cut -c1-600 gpurun_out/b20_reference.json

timeout 120 python tools/bench_vina.py > gpurun_out/vina_bench.json 2> gpurun_out/vina_bench.err || tail -3 gpurun_out/vina_bench.err; cut -c1-400 gpurun_out/vina_bench.json
for w in 3dbs 3dbs_x40; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-sustained > gpurun_out/f_$w.json 2> gpurun_out/f_$w.err || tail -5 gpurun_out/f_$w.err; python -c "
import json;d=json.load(open('gpurun_out/f_$w.json'));print('$w', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['final_lig_rmsd_A'], d['cpu_baseline']['value'])"; done
timeout 400 python bench.py --workload cfg3_16x40 --steps 20 --warmup 3 > gpurun_out/f_cfg3.json 2> gpurun_out/f_cfg3.err || tail -8 gpurun_out/f_cfg3.err; python -c "
import json;d=json.load(open('gpurun_out/f_cfg3.json'));print('cfg3', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity'], d.get('mdn'))"
for w in posebusters_256x40 revdock_512x40; do timeout 400 python bench.py --workload $w --complexes 16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/f_$w.json 2> gpurun_out/f_$w.err || tail -8 gpurun_out/f_$w.err; python -c "
import json;d=json.load(open('gpurun_out/f_$w.json'));print('$w x16', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('mdn'))"; done
K=11 timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/_gpu_dbg.py > gpurun_out/san_k11_memcheck.log 2>&1; tail -2 gpurun_out/san_k11_memcheck.log
K=11 timeout 300 compute-sanitizer --tool racecheck --print-limit 3 python tools/_gpu_dbg.py > gpurun_out/san_k11_racecheck.log 2>&1; tail -2 gpurun_out/san_k11_racecheck.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_vina.py -m gpu -q -k "score or rejects" > gpurun_out/san_vina_memcheck.log 2>&1; tail -3 gpurun_out/san_vina_memcheck.log

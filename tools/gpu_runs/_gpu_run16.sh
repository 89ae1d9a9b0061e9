K=6 timeout 90 python tools/_gpu_dbg.py > gpurun_out/smoke16.log 2>&1 || { tail -5 gpurun_out/smoke16.log; echo SMOKE_FAILED; exit 1; }
tail -2 gpurun_out/smoke16.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or golden_trajectory or sharded_sampling or bench_batch or full_size" 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b16_k6.json 2> gpurun_out/b16_k6.err || tail -5 gpurun_out/b16_k6.err
python -c "
import json;d=json.load(open('gpurun_out/b16_k6.json'));print(d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['final_lig_rmsd_A'], d['roofline'].get('sustained',{}).get('steps_per_s'), d['clocks'])"
timeout 300 python -m pytest tests/test_vina.py -m gpu -x -q 2>&1 | tail -4

K=11 timeout 60 python tools/_gpu_dbg.py > gpurun_out/smoke19.log 2>&1 || { tail -5 gpurun_out/smoke19.log; echo SMOKE_FAILED; exit 1; }
export B200DOCK_TEST_KERNELS=11
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or golden_trajectory or bench_batch" 2>&1 | tail -2
for k in 11 6; do timeout 300 python bench.py --conv-kernel $k --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b19_k$k.json 2> gpurun_out/b19_k$k.err || tail -5 gpurun_out/b19_k$k.err
python -c "
import json;d=json.load(open('gpurun_out/b19_k$k.json'));print($k, d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'], d['e2e']['value'], d['parity']['final_lig_rmsd_A'], d['roofline'].get('sustained',{}).get('steps_per_s'), d['clocks'])"; done
timeout 100 python tools/timeline.py 11 2>&1 | tee gpurun_out/timeline_k11.log | tail -12; timeout 100 python tools/timeline.py 6 2>&1 | tee gpurun_out/timeline_k6.log | tail -24

timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/_gpu_dbg.py > gpurun_out/san2.log 2>&1; tail -12 gpurun_out/san2.log
export B200DOCK_TEST_KERNELS=10
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/t3.log; tail -12 gpurun_out/t3.log
for k in 6 10; do timeout 200 python bench.py --conv-kernel $k --fast-kernel 0 --no-mdn --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b3_k$k.json 2> gpurun_out/b3_k$k.err; python -c "
import json;d=json.load(open('gpurun_out/b3_k$k.json'));print($k, d['roofline']['kernel_ms_per_step'], d['ms_per_step'])"; done

export B200DOCK_TEST_KERNELS=10
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_identical or golden" 2>&1 | tail -3 > gpurun_out/t6.log; tail -3 gpurun_out/t6.log
for k in 10 6; do timeout 200 python bench.py --conv-kernel $k --fast-kernel 0 --no-mdn --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b6_k$k.json 2> gpurun_out/b6_k$k.err; python -c "
import json;d=json.load(open('gpurun_out/b6_k$k.json'));print($k, d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'])"; done
timeout 600 ncu --set full --clock-control none -k regex:k_conv_v3 -c 7 -o gpurun_out/prof_v3c -f python bench.py --conv-kernel 10 --fast-kernel 0 --no-mdn --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu6.log 2>&1; tail -2 gpurun_out/ncu6.log

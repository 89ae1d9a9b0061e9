for k in 10 6 10; do timeout 200 python bench.py --conv-kernel $k --fast-kernel 0 --no-mdn --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b4_k$k.json 2> gpurun_out/b4_k$k.err; python -c "
import json;d=json.load(open('gpurun_out/b4_k$k.json'));print($k, d['roofline']['kernel_ms_per_step'], d['ms_per_step'], d['value'])"; done
timeout 600 ncu --set full --clock-control none -k regex:k_conv_v3 -c 8 -o gpurun_out/prof_v3 -f python bench.py --conv-kernel 10 --fast-kernel 0 --no-mdn --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu4.log 2>&1; tail -3 gpurun_out/ncu4.log

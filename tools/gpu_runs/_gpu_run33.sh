for flag in "" "--balance"; do
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload posebusters_256x40 --complexes 16 --steps 20 --warmup 2 --no-cpu-baseline $flag > gpurun_out/bal2$flag.json 2> gpurun_out/bal2$flag.err || tail -12 gpurun_out/bal2$flag.err
python -c "
import json;d=json.loads(open('gpurun_out/bal2$flag.json').read().strip().splitlines()[-1]);print('$flag', d['value'], d['e2e']['value'], d['config']['rank_device_ms'], d['config']['imbalance_max_over_mean'], d['mdn'])"
done

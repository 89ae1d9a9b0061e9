w=posebusters_256x40
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --workload $w --steps 20 --warmup 2 > gpurun_out/w8f_$w.json 2> gpurun_out/w8f_$w.err || tail -12 gpurun_out/w8f_$w.err
python -c "
import json;d=json.loads(open('gpurun_out/w8f_$w.json').read().strip().splitlines()[-1]);print('$w', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('mdn'), d['config']['rank_device_ms'], d['config']['imbalance_max_over_mean'], d['parity'])"

# BASELINE configs[3] and [4] at 8 GPUs through the product path (shard.run_sharded + device assembly + sampler + MDN + one all_gather)
for w in posebusters_256x40 revdock_512x40; do
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $w --steps 20 --warmup 2 > gpurun_out/w8_$w.json 2> gpurun_out/w8_$w.err || tail -12 gpurun_out/w8_$w.err
python -c "
import json;d=json.loads(open('gpurun_out/w8_$w.json').read().strip().splitlines()[-1]);print('$w', d['value'], d['e2e']['value'], d['ms_per_step'], d.get('mdn'), d['config']['rank_device_ms'], d['config']['imbalance_max_over_mean'], d['parity'])"
done

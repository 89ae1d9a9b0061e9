timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s2.json 2> gpurun_out/s2.err || tail -8 gpurun_out/s2.err
python -c "
import json;d=json.loads(open('gpurun_out/s2.json').read().strip().splitlines()[-1]);print('N=2', d['value'], d['e2e']['value'], d['ms_per_step'], d['parity']['final_lig_rmsd_A'], d['n_gpus'], d['gpu_launches'])"
B200DOCK_REF_BUDGET_S=20 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-300

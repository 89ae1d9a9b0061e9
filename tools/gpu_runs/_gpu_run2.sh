export B200DOCK_TEST_KERNELS=10
timeout 300 python - <<'PY' > gpurun_out/v3_smoke.log 2>&1
import torch, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from diffbindfr_b200 import synth, weights
from diffbindfr_b200.engine import Engine
from helpers import conditioning
sd = weights.random_state_dict(0)
b = synth.make_batch(**synth.WORKLOADS["tiny"], seed=3)
c = conditioning(b)
outs = {}
for k in (6, 10):
    eng = Engine(0, conv_kernel=k); eng.load_state_dict(sd)
    o = eng.score(b, c["t"], c["tr_sigma"], c["rot_score_norm"], c["tor_score_norm2"], c["sc_tor_score_norm2"])
    torch.cuda.synchronize()
    outs[k] = [x.cpu() for x in o]
    print("kernel", k, "ok", [float(x.abs().max()) for x in outs[k]], flush=True)
for a, r in zip(outs[10], outs[6]):
    print("equal", torch.equal(a, r), float((a - r).abs().max()))
PY
tail -5 gpurun_out/v3_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/t2.log; tail -8 gpurun_out/t2.log
for k in 6 10; do timeout 200 python bench.py --conv-kernel $k --fast-kernel 0 --no-mdn --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/b2_k$k.json 2> gpurun_out/b2_k$k.err; python -c "
import json;d=json.load(open('gpurun_out/b2_k$k.json'));print($k, d['roofline']['kernel_ms_per_step'], d['ms_per_step'])"; done

import torch, sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from diffbindfr_b200 import synth, weights
from diffbindfr_b200.engine import Engine
from helpers import conditioning
sd = weights.random_state_dict(0)
b = synth.make_batch(**synth.WORKLOADS["tiny"], seed=3)
c = conditioning(b)
k = int(os.environ.get("K", "10"))
eng = Engine(0, conv_kernel=k); eng.load_state_dict(sd)
if os.environ.get("LAYERS"): eng.debug_set(0, int(os.environ["LAYERS"]))
o = eng.score(b, c["t"], c["tr_sigma"], c["rot_score_norm"], c["tor_score_norm2"], c["sc_tor_score_norm2"])
torch.cuda.synchronize()
print("kernel", k, "ok", [float(x.abs().max()) for x in o], flush=True)

"""Where the fused conv kernel (modes 5 / 6) waits: cycle counters accumulated inside the kernel (b200dock_debug_set(1, 1)).
Run on a B200:  python tools/trace_waits.py [steps] [kernel]"""
import os, sys
_T = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffbindfr_b200", "libb200dock_trace.so")
assert os.path.exists(_T), "build the traced library first: python __graft_entry__.py --trace"
os.environ["B200DOCK_LIB"] = _T
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffbindfr_b200 import synth, weights, schedule
from diffbindfr_b200.engine import Engine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 6          # 5 (single CTA) or 6 (CTA pairs: the MMA warp lives in the even CTAs)
b = synth.make_batch(**synth.WORKLOADS["cfgA"], seed=0)
eng = Engine(0, conv_kernel=kernel)
eng.load_state_dict(weights.random_state_dict(0))
sch = schedule.make_schedule()[:steps]
B, n_tor, n_sc = b["num_graphs"], int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
noise = torch.randn(steps, 6 * B + n_tor + n_sc, generator=torch.Generator().manual_seed(1))
eng.sample(b, sch, noise)                      # warm-up
torch.cuda.synchronize()
eng.debug_set(1, 1)
eng.sample(b, sch, noise)
torch.cuda.synchronize()
t = eng.tap(7, dtype=np.int64).reshape(148, 32).astype(np.float64)
m = t.mean(0)
if kernel == 6:
    m[:8] = t[0::2, :8].mean(0)
names_m = ["x_full wait", "h_full wait", "d_empty wait", "b_full[0] wait", "b_full[1] wait", "b_full[2] wait", "-", "MMA warp total"]
names_e = ["a_empty wait", "xin gather+convert+store", "x1 gather + sh", "D1 (d_full) wait", "H1 conversion", "fold d_full waits", "fold compute", "epilogue total"]
print(f"mean cycles per CTA over {steps} steps (all conv launches); share of the warp's total")
for i, n in enumerate(names_m):
    if n != "-": print(f"  MMA warp  {n:28s} {m[i]:14.0f}  {100 * m[i] / m[7]:5.1f} %")
for i, n in enumerate(names_e):
    print(f"  epilogue  {n:28s} {m[8 + i]:14.0f}  {100 * m[8 + i] / m[15]:5.1f} %")

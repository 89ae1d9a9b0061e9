"""Per-unit timeline of the pair kernel on CTA 0 (traced build).  MMA warp: reaches the accumulator wait / issue begins / unit issued;
fold warps 0-3: accumulator seen full / released.  python tools/timeline.py  (on a B200)"""
import os, sys
_T = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "diffbindfr_b200", "libb200dock_trace.so")
os.environ["B200DOCK_LIB"] = _T
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffbindfr_b200 import synth, weights, schedule
from diffbindfr_b200.engine import Engine
b = synth.make_batch(**synth.WORKLOADS["cfgA"], seed=0)
eng = Engine(0, conv_kernel=6)
eng.load_state_dict(weights.random_state_dict(0))
sch = schedule.make_schedule()[10:12]
B, n_tor, n_sc = b["num_graphs"], int(b["tor_edge_mask"].sum()), int(b["sc_torsion_edge_mask"].sum())
noise = torch.randn(2, 6 * B + n_tor + n_sc, generator=torch.Generator().manual_seed(1))
eng.sample(b, sch, noise); torch.cuda.synchronize()
eng.debug_set(1, 1)
eng.sample(b, sch[:1], noise[:1]); torch.cuda.synchronize()
t = eng.tap(7, dtype=np.int64)
TL = 6144
base = 148 * 32
cnt = t[base:base + 16]
S = [t[base + 16 + i * TL: base + 16 + (i + 1) * TL][:min(int(cnt[i]), TL)] for i in range(11)]
print("events", cnt[:11])
tag = S[0] & 0xff
T = [(s & ~0xff).astype(np.int64) for s in S]
w2 = tag != 0                                   # W2 units (the W1 unit of a tile has tag 0)
A, Bg, Cc = T[0][w2], T[1][w2], T[2][w2]
n = min(len(A), *(len(T[i]) for i in range(3, 11)))
A, Bg, Cc = A[:n], Bg[:n], Cc[:n]
seen = np.stack([T[3 + q][:n] for q in range(4)]); rel = np.stack([T[7 + q][:n] for q in range(4)])
kind = S[7][:n] & 0xff
pc = lambda x: np.percentile(x, [5, 25, 50, 75, 95]).round().astype(int).tolist()
print("MMA warp: wait for the accumulator (reach -> free)      ", pc(Bg - A))
print("MMA warp: issue of one unit (free -> all issued)         ", pc(Cc - Bg))
print("MMA warp: issued -> reaches next wait                    ", pc(A[1:] - Cc[:-1]))
print("unit interval (issue begin to issue begin), ideal 2088   ", pc(np.diff(Bg)), "mean", float(np.diff(Bg)[np.diff(Bg) < 50000].mean()))
print("issued (commit queued) -> first fold warp sees it full   ", pc(seen.min(0) - Cc))
print("spread between fold warps seeing it full (max - min)     ", pc(seen.max(0) - seen.min(0)))
for k, nm in ((1, "w48"), (2, "w12")):
    m = kind == k
    print(nm, "fold hold time (seen -> released), warp 0", pc((rel[0] - seen[0])[m]), " slowest warp - fastest warp release", pc((rel.max(0) - rel.min(0))[m]))
# release of unit u (slowest local warp) -> MMA warp sees the accumulator of unit u+2 free
d = Bg[2:] - rel.max(0)[:-2]
print("slowest LOCAL release of unit u -> issue begin of unit u+2", pc(d), "(negative = the peer CTA was later)")
starts = np.where(tag == 0)[0]
print("tiles", len(starts))

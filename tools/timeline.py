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
KERNEL = int(sys.argv[1]) if len(sys.argv) > 1 else 6
eng = Engine(0, conv_kernel=KERNEL)
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
# tile structure from the MMA warp alone (works for every pair-kernel variant)
starts = np.where(tag == 0)[0]
ib = T[1]
rows = []
for s0, s1 in zip(starts[:-1], starts[1:]):
    nu = s1 - s0 - 1
    seg = ib[s0:s1 + 1]
    if (np.diff(seg) <= 0).any() or (np.diff(seg) > 400000).any(): continue
    rows.append((nu, seg[1] - seg[0], seg[2] - seg[1] if nu > 1 else 0, np.median(np.diff(seg[1:-1])) if nu > 2 else 0, seg[-1] - seg[-2], seg[-1] - seg[0]))
rows = np.array(rows, dtype=np.float64)
for nu in sorted(set(rows[:, 0].astype(int))):
    r = rows[rows[:, 0] == nu]
    print(f"tiles with {nu:2d} units: {len(r):3d}   W1 begin -> unit0 begin {np.median(r[:,1]):7.0f}   unit0 -> unit1 {np.median(r[:,2]):6.0f}   mid interval {np.median(r[:,3]):6.0f}"
          f"   last unit begin -> next W1 begin {np.median(r[:,4]):7.0f}   tile span {np.median(r[:,5]):8.0f}  ideal {(nu + 1) * 2088}")
# the first units of a tile in detail: per unit (reach accumulator wait -> free), (free -> all issued), (issued -> next unit reaches its wait)
det = {k: [] for k in range(-1, 4)}
for s0, s1 in zip(starts[:-1], starts[1:]):
    if s1 - s0 < 8: continue
    for k in range(-1, 4):
        i = s0 + 1 + k
        det[k].append((T[1][i] - T[0][i], T[2][i] - T[1][i], T[0][i + 1] - T[2][i], T[0][i] - T[2][i - 1] if i > 0 else 0))
for k in range(-1, 4):
    a = np.array(det[k], dtype=np.float64)
    print(("W1    " if k < 0 else f"unit {k}"), " before-wait gap", np.median(a[:, 3]), " accumulator wait", np.median(a[:, 0]), " issue", np.median(a[:, 1]))
tot = rows[:, 5].sum(); ideal = ((rows[:, 0] + 1) * 2088).sum()
print("tensor-busy share over these tiles (ideal / span):", ideal / tot)
if KERNEL != 6: sys.exit(0)
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

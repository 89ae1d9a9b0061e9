# Runs the reference's so3.py (copied verbatim to /tmp, the tree is read-only and the module
# writes a cache next to itself) with a stub ``io`` so its 1000-entry _exp_score_norms table
# can be stored as a golden fixture.
import sys, types, importlib.util, numpy as np, torch, time
pkg = types.ModuleType("gu"); pkg.__path__ = ["/tmp/refcopy/gu"]; sys.modules["gu"] = pkg
io = types.ModuleType("gu.io"); io._save_lmdb = lambda *a, **k: None; io._save = lambda *a, **k: None
sys.modules["gu.io"] = io; pkg.io = io
spec = importlib.util.spec_from_file_location("gu.so3_ref", "/tmp/refcopy/gu/so3_ref.py")
m = importlib.util.module_from_spec(spec); m.__package__ = "gu"
t = time.time(); spec.loader.exec_module(m); print("so3 tables built in", time.time() - t, "s")
torch.save(dict(exp_score_norms=torch.from_numpy(m._exp_score_norms)), "/root/repo/tests/golden/so3_exp_score_norms.pt")
print(m._exp_score_norms[:3], m._exp_score_norms[-3:])

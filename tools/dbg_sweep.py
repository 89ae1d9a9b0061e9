"""Timing experiments on the pair kernel (traced build, B200DOCK_DBG switches): conv ms/step with parts of the epilogue removed."""
import os, subprocess, sys, json
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for dbg in [0, 1, 2, 3, 4, 8, 16, 28, 31]:
    env = dict(os.environ, B200DOCK_LIB=os.path.join(root, "diffbindfr_b200", "libb200dock_trace.so"), B200DOCK_DBG=str(dbg))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline", "--no-sustained", "--steps", "20", "--warmup", "3"],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(dbg, "conv ms/step", round(d["roofline"]["kernel_ms_per_step"], 3), "step ms", round(d["ms_per_step"], 3), "sm_mhz", d["clocks"]["sm_mhz"], flush=True)
    except Exception as e:
        print(dbg, "failed", r.stderr[-400:], flush=True)

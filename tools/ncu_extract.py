"""Extract per-launch duration / DRAM traffic / tensor-pipe activity from an ncu --set full report.

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep profiles/name.json [profiles/name.md]
"""
import csv, json, subprocess, sys

COLS = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_bytes",
        "launch__registers_per_thread": "regs", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
        "smsp__inst_executed.sum": "warp_insts"}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "msecond": 1e-3,
         "nsecond": 1e-9, "second": 1}


def main(rep, out_json, out_md=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[ix["Kernel Name"]].split("(")[0]}
        for c, name in COLS.items():
            if c in ix:
                v = float(r[ix[c]].replace(",", "") or 0)
                d[name] = v * SCALE.get(units[ix[c]], 1)
        launches.append(d)
    tot = {k: sum(l.get(k, 0) for l in launches) for k in ("duration", "dram_read", "dram_write", "l2_to_sm_bytes")}
    out = {"report": rep, "launches": launches, "sum": tot,
           "dram_bytes_per_launch": (tot["dram_read"] + tot["dram_write"]) / max(len(launches), 1)}
    json.dump(out, open(out_json, "w"), indent=1)
    if out_md:
        with open(out_md, "w") as f:
            f.write(f"# ncu --set full --clock-control none, `{rep}`\n\n| # | kernel | ms | DRAM read MB | DRAM write MB | tensor pipe % | L2 hit % | L2->SM GB | regs |\n|---|---|---:|---:|---:|---:|---:|---:|---:|\n")
            for i, l in enumerate(launches):
                f.write(f"| {i} | {l['kernel']} | {l['duration']*1e3:.3f} | {l['dram_read']/1e6:.1f} | {l['dram_write']/1e6:.1f} | "
                        f"{l.get('tensor_pipe_pct', 0):.1f} | {l.get('l2_hit_pct', 0):.1f} | {l.get('l2_to_sm_bytes', 0)/1e9:.2f} | {int(l.get('regs', 0))} |\n")
            f.write(f"\nsum: {tot['duration']*1e3:.3f} ms, DRAM {tot['dram_read']/1e6:.0f} MB read + {tot['dram_write']/1e6:.0f} MB written "
                    f"({out['dram_bytes_per_launch']/1e6:.0f} MB per launch)\n")
    print(json.dumps(out["sum"]), out["dram_bytes_per_launch"])


if __name__ == "__main__":
    main(*sys.argv[1:])

"""Turns the raw ncu outputs a gpurun call brings back (gpurun_out/) into the small tracked
summaries under profiles/:
  launches : ncu --metrics gpu__time_duration.sum --csv launch list  -> per-kernel totals + shares
  full     : ncu --set full .ncu-rep (read with `ncu -i ... --page raw --csv`) -> key metrics JSON
usage: python tools/ncu_summary.py launches <launches.csv> <out.csv> [note]
       python tools/ncu_summary.py full <file.ncu-rep> <out.json> [traffic.json]"""
import csv
import json
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("rmi::<unnamed>::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    name = name.replace("unsigned long long", "u64").replace("unsigned int", "u32")
    return name.strip()


def launches(src, dst, note=""):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
        name = short(r[ki])
        if name.startswith("k_leaf"):      # the bulk kernel and the one-warp-per-block long-leaf kernel share a name
            name += " grid" + r[hdr.index("Grid Size")].replace(" ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    tot = sum(v[1] for v in ours.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({note})\n# `ncu --metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised: compare SHARES\n")
        f.write("kernel,launches,total_ms,avg_ms,share_of_rmi_kernels\n")
        for k, (c, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{c},{ms:.4f},{ms / c:.4f},{ms / tot:.3f}\n")
        f.write("# other kernels in the same process (torch sort/randint of the synthetic keys, NCCL, ...: not part of a build):\n")
        for k, (c, ms) in sorted(((k, v) for k, v in agg.items() if k not in ours), key=lambda kv: -kv[1][1])[:6]:
            f.write(f"# {k[:90]},{c},{ms:.3f}\n")


def full(rep, dst, traffic=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[hdr.index("Kernel Name")])}
        for m in KEEP:
            if m in hdr:
                d[m] = f"{r[hdr.index(m)]} {units[hdr.index(m)]}".strip()
        res.append(d)
    json.dump(res, open(dst, "w"), indent=1)
    if traffic and res:
        def to_bytes(s):
            v, u = s.split()[0], (s.split() + [""])[1]
            return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        # the bulk leaf kernel of ONE build may be several launches (slices of the block range): sum them
        bulk = [r for r in res if not r.get("launch__grid_size", "").strip().startswith("16")]
        rd = sum(to_bytes(r["dram__bytes_read.sum"]) for r in bulk)
        wr = sum(to_bytes(r["dram__bytes_write.sum"]) for r in bulk)
        def to_us(s):
            v, u = s.split()[0], (s.split() + [""])[1]
            return float(v.replace(",", "")) * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)
        json.dump({"kernel": bulk[0]["kernel"], "source": dst, "launches_summed": len(bulk),
                   "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
                   "gpu_time_us_sum": sum(to_us(r["gpu__time_duration.sum"]) for r in bulk),
                   "lts_hit_rate": [r.get("lts__t_sector_hit_rate.pct") for r in bulk]}, open(traffic, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:])

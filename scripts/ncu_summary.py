"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed):

    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/<tag>_launches_summary.csv "<command>"
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/<tag>_top_kernels_ncu_full.json "<command>"
"""
import collections, csv, json, re, subprocess, sys


def clean(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")


def launches(src, dst, cmd):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[h], rows[h + 2:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v, u = float(r[vi].replace(",", "")), r[ui]
        us = v / 1e3 if u.startswith("ns") else (v if u.startswith("us") else v * 1e3)
        a = agg.setdefault(clean(r[ki])[:110], [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
    out = [f"# ncu launch list: `ncu --metrics gpu__time_duration.sum --clock-control none {cmd}`",
           "# every launch of the run (warm-up steps included); durations are cold-cache and serialised: compare SHARES with "
           "bench.py's live CUDA-event shares", f"# {len(data)} launches, {tot / 1e3:.3f} ms total", "kernel,launches,total_us,share"]
    out += [f"{k},{n},{t:.1f},{t / tot:.4f}" for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


def full(src, dst, cmd):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = {"gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dr", "dram__bytes_write.sum": "dw",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "regs",
            "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
            "launch__grid_size": "grid", "launch__block_size": "block"}
    idx = {v: hdr.index(k) for k, v in want.items() if k in hdr}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}

    def num(r, k):
        try:
            return float(r[idx[k]].replace(",", "")) * scale.get(units[idx[k]], 1)
        except (ValueError, KeyError):
            return None

    ki = hdr.index("Kernel Name")
    out = []
    for r in data:
        t, dr, dw = num(r, "time"), num(r, "dr"), num(r, "dw")
        d = {"kernel": clean(r[ki]), "time_ms": round(t * 1e3, 4), "dram_read_GB": round(dr / 1e9, 4),
             "dram_write_GB": round(dw / 1e9, 4), "dram_bytes_per_launch": int(dr + dw), "dram_GBps": round((dr + dw) / t / 1e9, 1)}
        for k in ("dram_pct_of_peak", "l2_pct_of_peak", "sm_pct", "tensor_pipe_pct", "warps_active_pct", "regs", "l1_hit_pct",
                  "l2_hit_pct", "grid", "block"):
            v = num(r, k)
            d[k] = None if v is None else round(v, 3)
        out.append(d)
    json.dump({"source": f"ncu --set full --clock-control none --import-source on, {cmd} (one replayed launch each, cold cache, "
                         "serialised); units: ms, GB, bytes, GB/s, %", "kernels": out}, open(dst, "w"), indent=1)
    for d in out:
        print(d["kernel"][:44].ljust(44), d["time_ms"], d["dram_read_GB"], d["dram_write_GB"], d["dram_GBps"], d["tensor_pipe_pct"],
              d["l2_pct_of_peak"], d["l2_hit_pct"])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](*sys.argv[2:5])

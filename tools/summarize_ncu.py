"""Summarise gpurun_out ncu artefacts into profiles/ (run here, no GPU needed).
usage: summarize_ncu.py launches <launches.csv> <out.md> | full <report.ncu-rep> <out.md>"""
import csv, io, subprocess, sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def launches(path, out):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("covo::", "")
        us = float(r[iv].replace(",", "")) / 1e3 if "us" not in hdr[iv] else float(r[iv])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    unit_ns = True
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; serialised, cold-cache: compare SHARES)\n\n")
        f.write(f"source: `{path}`; {sum(v[0] for v in agg.values())} launches, {tot:.0f} us\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {100*us/tot:.1f} % |\n")
    print(open(out).read())


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{path}` (--clock-control none)\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## {d.get('Kernel Name','?')[:90]}  grid {d.get('Grid Size','')} block {d.get('Block Size','')}\n")
            for h, u, v in zip(hdr, units, r):
                if h in KEYS:
                    f.write(f"- {h} [{u}] = {v}\n")
            st = {h.split('stalled_')[1]: float(v) for h, v in zip(hdr, r) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h}
            tot = sum(st.values()) or 1
            f.write("- warp-stall samples: " + ", ".join(f"{k} {100*v/tot:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]) + "\n")
            try:
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
                tr = sum(float(d[k].replace(",", "")) * scale[units[hdr.index(k)]]
                         for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                f.write(f"- DRAM traffic = read + write = {tr:.1f} Mbyte\n")
            except Exception:
                pass
            f.write("\n")
    print(open(out).read()[:6000])


def dram(path, out):
    """Per-launch time + DRAM bytes of one vocoder forward (ncu --metrics gpu__time_duration.sum,dram__bytes_*): by kernel."""
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    iid, ik, im, iu, iv = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    per = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        d = per.setdefault(r[iid], {"name": r[ik].split("(")[0].replace("void ", "").replace("covo::", "")})
        d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    launches_ = list(per.values())
    # the capture holds warm-up forwards too: keep the LAST forward (from the last mel_to_tc_kernel on)
    starts = [i for i, d in enumerate(launches_) if d["name"].startswith("mel_to_tc")]
    if starts:
        launches_ = launches_[starts[-1]:]
    agg = OrderedDict()
    for d in launches_:
        a = agg.setdefault(d["name"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    t_us = sum(a[1] for a in agg.values())
    by = sum(a[2] for a in agg.values())
    with open(out, "w") as f:
        f.write("# HiFi-GAN forward at [8, 80, 1500]: ncu per-launch time and DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum,\n"
                "# --clock-control none; launches serialised by the profiler, so times are per-kernel, not the graph's)\n\n")
        f.write(f"source: `{path}`; last forward of the capture: {len(launches_)} launches, {t_us:.0f} us, {by/1e9:.2f} GB of DRAM traffic "
                f"= {by/1e3/t_us:.0f} GB/s averaged over the kernels' own time\n\n| kernel | launches | us | DRAM GB | GB/s |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {b/1e9:.3f} | {b/1e3/max(us,1e-9):.0f} |\n")
        f.write("\nper launch (time order):\n\n| # | kernel | us | DRAM MB | GB/s |\n|---:|---|---:|---:|---:|\n")
        for i, d in enumerate(launches_):
            b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            us = d.get("gpu__time_duration.sum", 0.0)
            f.write(f"| {i} | `{d['name']}` | {us:.1f} | {b/1e6:.1f} | {b/1e3/max(us,1e-9):.0f} |\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "full": full, "dram": dram}[sys.argv[1]](sys.argv[2], sys.argv[3])

"""Turn gpurun_out/launches_*.csv + prof_*.ncu-rep into the committed summaries under profiles/."""
import collections, csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PREC = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
ROUND = sys.argv[2] if len(sys.argv) > 2 else "r1"


def launches():
    path = os.path.join(OUT, f"launches_{PREC}.csv")
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:70]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    out = [f"# ncu launch list, bench.py --precision {PREC} --batch 64 (one train step after 3 warm-up steps)", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -s <3 warm-up steps> -c <one step>` (profiles/*.sh).",
           "Per-launch times are cold-cache and serialised: read the SHARES.", "",
           f"{sum(cnt.values())} launches, {T/1e3:.2f} ms of device time", "",
           "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    ours = 0.0
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:30]:
        out.append(f"| `{k}` | {cnt[k]} | {v:.0f} | {100*v/T:.1f}% | {v/cnt[k]:.1f} |")
    ours = sum(v for k, v in tot.items() if k.startswith("crk::"))
    out += ["", f"library kernels (`crk::*`): {100*ours/T:.1f}% of device time; the rest is torch glue "
            "(cat / embedding / dropout-mask RNG / NCCL-free elementwise) ."]
    return "\n".join(out)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct",
        "dram__cycles_active.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg ",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warps_issue_stalled", "sm__pipe_tensor_subpipe"]


def full(kernel):
    rep = os.path.join(OUT, f"prof_{kernel}_{PREC}.ncu-rep")
    if not os.path.exists(rep):
        return f"## {kernel}\n(no capture)\n"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = [f"## {kernel} (`ncu --set full --clock-control none --import-source on`, 1 launch, {PREC})", "",
           f"kernel: `{vals[hdr.index('Kernel Name')][:110]}`", "", "| metric | value | unit |", "|---|---:|---|"]
    sel = []
    for i, h in enumerate(hdr):
        if any(h.startswith(k) or k in h for k in KEYS):
            if "issue_stalled" in h and "ratio" not in h:
                continue
            sel.append((h, vals[i], units[i]))
    for h, v, u in sorted(sel):
        out.append(f"| {h} | {v} | {u} |")
    return "\n".join(out) + "\n"


with open(os.path.join(ROOT, "profiles", f"launches_{ROUND}_{PREC}.md"), "w") as f:
    f.write(launches() + "\n")
KERNELS = ["k_resblock_fwd_tc", "k_conv_tc", "k_wgrad_tc_raw", "k_wgrad_tc", "k_vq_argmin_tc"]
if any(os.path.exists(os.path.join(OUT, f"prof_{k}_{PREC}.ncu-rep")) for k in KERNELS):   # keep the last summary otherwise
    with open(os.path.join(ROOT, "profiles", f"ncu_full_{ROUND}_{PREC}.md"), "w") as f:
        f.write(f"# ncu --set full summaries ({ROUND}, {PREC}); .ncu-rep files stay in gpurun_out/ (scratch)\n\n")
        for k in KERNELS:
            f.write(full(k) + "\n")
print(open(os.path.join(ROOT, "profiles", f"launches_{ROUND}_{PREC}.md")).read()[:3000])

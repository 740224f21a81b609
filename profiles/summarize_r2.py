"""Round 2: turn the CSV exports of profiles/call_r2_17.sh (gpurun_out/launches_r2_*.csv, ncu_r2_<kernel>_{raw,source}.csv)
into the committed summaries profiles/launches_r2_tf32x3.md, profiles/ncu_full_r2_tf32x3.md, profiles/ncu_traffic_r2.json."""
import collections, csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PREC = "tf32x3"
csv.field_size_limit(1 << 30)


def launches():
    path = os.path.join(OUT, f"launches_r2_{PREC}.csv")
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
    out = [f"# ncu launch list, round 2, bench.py --precision {PREC} --batch 64 (one train step after 3 warm-up steps)", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -s <3 warm-up steps> -c 1000` (profiles/call_r2_17.sh).",
           "Per-launch times are cold-cache and serialised: read the SHARES.", "",
           f"{sum(cnt.values())} launches, {T/1e3:.2f} ms of device time", "",
           "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:32]:
        out.append(f"| `{k}` | {cnt[k]} | {v:.0f} | {100*v/T:.1f}% | {v/cnt[k]:.1f} |")
    ours = sum(v for k, v in tot.items() if k.startswith("crk::"))
    out += ["", f"library kernels (`crk::*`): {100*ours/T:.1f}% of device time; the rest is torch glue "
            "(cat / embedding / dropout-mask RNG / elementwise)."]
    return "\n".join(out)


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct",
        "dram__cycles_active.avg.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg ",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warps_issue_stalled", "sm__pipe_tensor_subpipe",
        "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct", "launch__occupancy_per_block"]


def raw(kernel):
    path = os.path.join(OUT, f"ncu_r2_{kernel}_raw.csv")
    if not os.path.exists(path) or os.path.getsize(path) < 1000:
        return None
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[2]


def full(kernel):
    r = raw(kernel)
    if r is None:
        return f"## {kernel}\n(no capture)\n", None
    hdr, units, vals = r
    name = vals[hdr.index("Kernel Name")]
    out = [f"## {kernel} (`ncu --set full --clock-control none --import-source on`, 1 launch, {PREC})", "",
           f"kernel: `{name[:120]}`", "", "| metric | value | unit |", "|---|---:|---|"]
    sel = []
    for i, h in enumerate(hdr):
        if any(h.startswith(k) or k in h for k in KEYS):
            if "issue_stalled" in h and "ratio" not in h:
                continue
            sel.append((h, vals[i], units[i]))
    for h, v, u in sorted(sel):
        out.append(f"| {h} | {v} | {u} |")
    g = lambda key: next((float(vals[i].replace(",", "")) for i, h in enumerate(hdr) if h == key and vals[i] not in ("", "n/a")), None)
    def conv(key):   # value in bytes / us
        i = hdr.index(key)
        v = float(vals[i].replace(",", ""))
        u = units[i]
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)
        return v * mul
    tensor = next((float(vals[i]) for i, h in enumerate(hdr) if h.startswith("sm__pipe_tensor_cycles_active") and "pct_of_peak_sustained_elapsed" in h and vals[i] not in ("", "n/a")), None)
    info = {"kernel": name[:120], "grid": g("launch__grid_size"), "duration_us_under_ncu": conv("gpu__time_duration.sum"),
            "dram_bytes_read": conv("dram__bytes_read.sum"), "dram_bytes_write": conv("dram__bytes_write.sum"),
            "tensor_pipe_active_pct_elapsed": tensor, "registers": g("launch__registers_per_thread"),
            "ctas_per_sm_limit_smem": g("launch__occupancy_limit_shared_mem")}
    # hottest source lines
    sp = os.path.join(OUT, f"ncu_r2_{kernel}_source.csv")
    if os.path.exists(sp) and os.path.getsize(sp) > 100:
        rows = list(csv.reader(open(sp)))
        h2 = rows[1]
        si = h2.index("# Samples")
        data = [r for r in rows[2:] if len(r) > si and r[si].isdigit()]
        tot = sum(int(r[si]) for r in data) or 1
        top = sorted(data, key=lambda r: -int(r[si]))[:12]
        out += ["", f"hottest SASS instructions ({tot} stall samples):", "", "| samples | share | instruction |", "|---:|---:|---|"]
        for r in top:
            out.append(f"| {r[si]} | {100*int(r[si])/tot:.1f}% | `{r[1].strip()[:80]}` |")
    return "\n".join(out) + "\n", info


if __name__ == "__main__":
    with open(os.path.join(ROOT, "profiles", f"launches_r2_{PREC}.md"), "w") as f:
        f.write(launches() + "\n")
    kernels = ["k_resblock_fwd_tc2", "k_conv_tc_dgrad", "k_conv_tc_gate", "k_wgrad_tc_raw", "k_wgrad_tc", "k_vq_argmin_tf32", "k_logmel"]
    traffic = {"source": "ncu --set full --clock-control none, one launch each inside bench.py --precision tf32x3 (profiles/call_r2_17.sh); B200, 64 x 500 frames", "kernels": {}}
    with open(os.path.join(ROOT, "profiles", f"ncu_full_r2_{PREC}.md"), "w") as f:
        f.write(f"# ncu --set full summaries (round 2, {PREC}); reports were exported as CSV on the GPU box (profiles/call_r2_17.sh)\n\n")
        for k in kernels:
            text, info = full(k)
            f.write(text + "\n")
            if info:
                traffic["kernels"][k] = info
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic_r2.json"), "w"), indent=1)
    print(open(os.path.join(ROOT, "profiles", f"launches_r2_{PREC}.md")).read()[:2500])
    print(json.dumps(traffic, indent=1)[:3000])

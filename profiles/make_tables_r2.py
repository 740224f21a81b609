"""Round 2: gpurun_out/r2_sweep.jsonl, r2_scale_*_n*.json, r2_sanitizer_*.log -> profiles/sweep_r2.md, scaling_r2.md, sanitizer_r2.md
(+ copies of the bench lines they summarise)."""
import glob, json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PRO = os.path.join(ROOT, "profiles")


def last_json(path):
    for l in reversed(open(path).read().strip().splitlines()):
        if l.startswith("{"):
            return json.loads(l)
    raise ValueError(path)


def sweep():
    p = os.path.join(OUT, "r2_sweep.jsonl")
    if not os.path.exists(p):
        return
    rows = []
    for l in open(p):
        try:
            rows.append(json.loads(l))
        except Exception:
            pass
    out = ["# Round 2: workload variants and the BASELINE config-5 sweep subset on ONE B200 (`profiles/call_r2_19.sh`)", "",
           "`bench.py --steps 8 --warmup 3`, synthetic VCC2020-shaped batches, frames/s = B x T / step time (device-timed, batch resident in HBM).",
           "", "| trainer | utts/GPU | frames | arithmetic | CUDA graph | ms/step | frames/s | launches/step |", "|---|---:|---:|---|---|---:|---:|---:|"]
    for d in rows:
        c = d["config"]
        out.append(f"| {c['trainer']} | {c['batch_per_gpu']} | {c['frames']} | {d['dtype']} | {'yes' if c['cuda_graph'] else 'no'} | "
                   f"{d['ms_per_step']:.2f} | {d['value']:,.0f} | {d['gpu_launches'] / max(d['steps'], 1):.0f} |")
    out += ["", "Configs of BASELINE.json: (2) vqvae 16 utts = row 1; (3) lsgan, 64 utts on one GPU = the default bench line, 8 utts/GPU = its per-GPU "
            "share on 8 GPUs (rows `lsgan 8 500`); (4) cyclegan; (5) the T x B grid (cells above 140 000 frames per step skipped: "
            "saved activations of four networks at 0.7 GB per 32 000 frames per generator forward).", ""]
    open(os.path.join(PRO, "sweep_r2.md"), "w").write("\n".join(out))
    json.dump(rows, open(os.path.join(PRO, "bench_r2_sweep.json"), "w"), indent=0)


def scaling():
    files = sorted(glob.glob(os.path.join(OUT, "r2_scale_*_n*.json")))
    tab = {}
    for f in files:
        m = re.match(r"r2_scale_(\w+?)_n(\d+)\.json", os.path.basename(f))
        try:
            d = last_json(f)
        except Exception:
            continue
        tab[(m.group(1), int(m.group(2)))] = d
    one = {}
    # N = 1 of the SAME build as the multi-GPU runs (they were taken before the last single-GPU optimisations of the round)
    for name in ("r2_bench_22_m256.json", "r2_bench_21.json", "r2_bench_18.json"):
        p = os.path.join(OUT, name)
        if os.path.exists(p):
            one["weak"] = last_json(p)
            break
    if not tab:
        return
    out = ["# Round 2: scaling on one 8 x B200 node (`profiles/call_r2_20.sh`; one process per GPU, NCCL)", "",
           "weak = 64 utterances per GPU (the driver's SCALE run); strong = BASELINE config 3 as written: global batch 64 utterances split over the ranks.",
           "Device-timed, max over ranks, 20 timed steps after 5 warm-up steps.  The N = 8 runs and their N = 1 baseline are one build (20.9 ms per step on one GPU);",
           "the N = 2 points are an earlier build of the round (22.4 ms on one GPU); the final build runs the one-GPU step in 19.0 ms.", "",
           "| mode | GPUs | utts/GPU | ms/step | frames/s (all GPUs) | efficiency |", "|---|---:|---:|---:|---:|---:|"]
    base = {}
    for mode in ("weak", "strong", "stronggraph"):
        for n in (1, 2, 4, 8):
            d = tab.get((mode, n)) or (one.get(mode) if n == 1 else None)
            if d is None and n == 1 and mode.startswith("strong"):
                d = one.get("weak")         # global batch 64 on one GPU is the default line
            if d is None:
                continue
            if n == 1:
                base[mode] = d["value"]
            b = base.get(mode)
            eff = ""
            if b:
                eff = f"{d['value'] / (b * n):.3f}" if mode == "weak" else f"{d['value'] / b / n:.3f}"
            out.append(f"| {mode} | {n} | {d['config']['batch_per_gpu']} | {d['ms_per_step']:.2f} | {d['value']:,.0f} | {eff} |")
    out.append("")
    open(os.path.join(PRO, "scaling_r2.md"), "w").write("\n".join(out))
    json.dump({f"{k[0]}_n{k[1]}": v for k, v in tab.items()}, open(os.path.join(PRO, "bench_r2_scaling.json"), "w"), indent=0)


def sanitizer():
    out = ["# Round 2: compute-sanitizer over the GPU test suite (`profiles/call_r2_19.sh`)", ""]
    found = False
    for tool in ("memcheck", "racecheck"):
        p = os.path.join(OUT, f"r2_sanitizer_{tool}.log")
        if os.path.exists(p):
            found = True
            out += [f"## {tool}", "", "```"] + open(p).read().strip().splitlines()[-14:] + ["```", ""]
    if found:
        open(os.path.join(PRO, "sanitizer_r2.md"), "w").write("\n".join(out))


if __name__ == "__main__":
    sweep(); scaling(); sanitizer()
    for f in ("sweep_r2.md", "scaling_r2.md", "sanitizer_r2.md"):
        p = os.path.join(PRO, f)
        if os.path.exists(p):
            print(open(p).read())

"""Turn the raw artefacts a GPU run left in gpurun_out/ into the summaries committed under profiles/:
timeline_<tag>.json (scripts/timeline_step.py) -> profiles/r01_timeline_{train,infer}_by_kernel.txt,
launches_final.csv (ncu --metrics gpu__time_duration.sum launch list) -> profiles/r01_train_step_launches_by_kernel.txt."""
import collections, csv, json, os, re, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    return re.sub(r"\(.*$", "", n)[:110]


for tag, out, what in (("final", "r01_timeline_train_by_kernel.txt", "training"), ("final_infer", "r01_timeline_infer_by_kernel.txt", "inference")):
    src = os.path.join(OUT, f"timeline_{tag}.json")
    if not os.path.exists(src):
        continue
    rows = json.load(open(src))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, s, d in rows:
        a = agg[short(n)]; a[0] += 1; a[1] += d
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, out), "w") as f:
        f.write(f"# One warm replay of the graphed MViTv2-S {what} step (8 clips, bf16, B200): CUPTI kernel records via\n"
                "# scripts/timeline_step.py (torch.profiler).  Diagnostic only (a run under a profiler is not a bench value).\n"
                "# Durations are warm and in-graph, unlike the cold, serialised ncu launch list; gaps between kernels inside the\n"
                "# graph are <= 1 us, so the step is the sum of these durations.\n")
        f.write(f"# device activities {len(rows)}  busy {tot / 1e3:.3f} ms\n{'total us':>10s} {'share':>6s} {'n':>5s} {'avg us':>8s}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:8.1f}  {k}\n")
src = os.path.join(OUT, "launches_final.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(PROF, "r01_train_step_launches.csv"))
    r = list(csv.reader([l for l in open(src) if not l.startswith("==")]))
    h = r[0]
    ik, iv, im, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r[1:]:
        if len(row) > iv and row[im] == "gpu__time_duration.sum":
            a = agg[short(row[ik])]; a[0] += 1; a[1] += float(row[iv].replace(",", "")) * (1e-3 if row[iu].startswith("n") else 1.0)
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(PROF, "r01_train_step_launches_by_kernel.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1300 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline\n"
                "# (1300 launches of the eager training step, about 1.4 steps; cold-cache, serialised per-launch times: shares, not absolutes)\n"
                f"{'total us':>10s} {'share':>6s} {'n':>5s} {'avg us':>8s}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d} {v[1] / v[0]:8.1f}  {k}\n")
for name in ("r01_bench_train.json", "r01_bench_infer.json", "r01_bench_train_mvitv2_b.json"):
    if os.path.exists(os.path.join(OUT, name)):
        shutil.copy(os.path.join(OUT, name), os.path.join(PROF, name))
print("profiles/ refreshed")

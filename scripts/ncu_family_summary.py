"""Aggregate an ncu launch list (--csv, metrics gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum
[lts__t_bytes.sum sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active]) by kernel family.
    python scripts/ncu_family_summary.py gpurun_out/step_metrics.csv > profiles/r02_train_step_dram_by_family.txt"""
import collections, csv, re, sys

FAMILIES = [("attention_fwd", r"attn_tc_fwd_kernel|attn_fwd_kernel"), ("attention_bwd", r"attn_bwd_|attn_delta|cast_rows96"),
            ("gemm_tcgen05", r"gemm_tc_kernel"), ("gemm_ffma", r"gemm_simt"), ("pooling_fwd", r"pool_ws_fwd|pool_tma_kernel<.*\(int\)0|pool_fwd"),
            ("pooling_bwd", r"pool_|reduce_jobs"), ("relpos", r"relpos"), ("colsum_cast", r"colsum"), ("layernorm", r"layernorm"),
            ("partial_reduce", r"reduce_partials"), ("maxpool_skip", r"maxpool"), ("adamw", r"adamw|sumsq|clip"), ("im2col", r"im2col"),
            ("head", r"head_")]


def family(name):
    for fam, pat in FAMILIES:
        if re.search(pat, name):
            return fam
    return "torch / other"


rows = list(csv.reader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
h = rows[0]
ik, iv, im, iu, iid = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit"), h.index("ID")
per = collections.defaultdict(dict)
names = {}
for r in rows[1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(",", "") or 0)
    u = r[iu]
    if r[im] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    elif "bytes" in r[im]:
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    per[r[iid]][r[im]] = v
    names[r[iid]] = r[ik]
agg = collections.defaultdict(lambda: collections.defaultdict(float))
for i, m in per.items():
    a = agg[family(names[i])]
    a["n"] += 1
    a["us"] += m.get("gpu__time_duration.sum", 0.0)
    a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    a["l2"] += m.get("lts__t_bytes.sum", 0.0)
    a["tensor"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * m.get("gpu__time_duration.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
print(f"launches {len(per)} (eager launches of the training step, ncu serialised: cold caches, shares not absolutes)  total {tot:.0f} us")
for fam, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{fam:16s} n={int(a['n']):4d} {a['us']:9.1f} us ({100 * a['us'] / tot:5.1f}%)  dram {a['dram']:9.1f} MB -> {a['dram'] / max(a['us'], 1e-9) * 1e3:7.1f} GB/s"
          f"  ({a['dram'] / max(a['n'], 1):7.1f} MB per launch)" + (f"   L2 {a['l2']:9.1f} MB" if a["l2"] else "") +
          (f"   tensor-active {a['tensor'] / max(a['us'], 1e-9):5.1f}%" if a["tensor"] else ""))

"""Build profiles/<tag>_summary.md from the ncu CSV exports and the bench JSON of the same code state."""
import collections
import csv
import json
import re
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01_v5"
P = "profiles/"
bench = json.loads(open(P + f"{tag}_bench.json").readline())
out = []
w = out.append
rf, e2e, cb = bench["roofline"], bench["e2e"], bench.get("cpu_baseline") or {}
w(f"# {tag} — profile summary (B200, second-order MAML K=1, S=Q=4, 128 phonemes -> 864 frames, dropout ON)\n")
w(f"bench.py (CUDA-graph replay, clocks {bench['clocks']}): **{bench['ms_per_step']:.2f} ms / outer step = "
  f"{bench['value']:.0f} mel-frames/s**; end to end through MetaSystem.training_step + optimizer_step with host batches: "
  f"{e2e['value']:.0f} mel-frames/s ({e2e['ms_per_step']:.2f} ms/step, runs {['%.2f' % x for x in e2e['runs_ms_per_step']]}); "
  f"CPU oracle on the same box: {cb.get('value', float('nan')):.0f} mel-frames/s ({cb.get('cores')} threads).\n")
w(f"Dominant kernel `{rf['kernel'].split(' ')[0]}`: {rf['launches_per_step']} launches/step, {100 * rf['share_of_step']:.0f}% of the step, "
  f"**{rf['achieved']:.1f} algorithmic TFLOP/s = {100 * rf['frac']:.1f}% of the measured BURST bf16 peak ({rf['peak']} TF/s; the kernel is timed alone)**; "
  f"it issues {rf.get('tensor_pipe_work_multiplier', 3)} MMA pass(es) per algorithmic FLOP => {100 * rf['frac_of_issued_mma']:.1f}% of peak as issued tensor work.  "
  f"All GEMM kernels: {rf['all_gemm']['ms_per_step']:.2f} ms/step of kernel time ({rf['all_gemm']['tflops']:.1f} TFLOP/s; they overlap "
  f"across the main / side / branch streams, so the sum exceeds their share of the step).\n")
w("| tensor-core kernel variant | launches/step | ms/step (graph replay of that variant's launches, CUDA events) | algorithmic GFLOP | TFLOP/s |")
w("|---|---|---|---|---|")
for v in rf["all_gemm_kernels"]:
    w(f"| `{v['kernel']}` | {v['launches']} | {v['ms_per_step']:.3f} | {v['algorithmic_gflop']:.1f} | {v['tflops']:.1f} |")

# ---- launch list
lines = [l for l in open(P + f"{tag}_ncu_launches.csv") if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ik, iv, im, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
agg, n = collections.OrderedDict(), 0
for row in r:
    if len(row) <= iv or row[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row[ik]).replace("<unnamed>::", "").replace("void ", "")
    v = float(row[iv].replace(",", ""))
    v = v / 1e3 if row[iu] == "ns" else (v * 1e3 if row[iu] == "ms" else v)
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += v
    n += 1
tot = sum(v for _, v in agg.values())
w(f"\n## ncu launch list of one eager outer step (`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`)\n")
w(f"Serialised, cold-cache, un-boosted clocks: compare shares.  CSV: `profiles/{tag}_ncu_launches.csv`.\n")
w(f"launches {n}, summed device time {tot / 1e3:.2f} ms\n")
w("| kernel | launches | ms | share | avg us |")
w("|---|---|---|---|---|")
gemm = 0.0
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    w(f"| `{k[:70]}` | {c} | {v / 1e3:.3f} | {100 * v / tot:.1f}% | {v / c:.1f} |")
for k, (c, v) in agg.items():
    if "mtts_gemm" in k:
        gemm += v
w(f"\nShare of the GEMM kernels in the ncu list: {100 * gemm / tot:.1f}% (bench.py's CUDA-event share of the dominant variant: "
  f"{100 * rf['share_of_step']:.0f}%).\n")

# ---- full-set capture
f = csv.reader(open(P + f"{tag}_ncu_gemm_raw.csv"))
hdr = next(f)
next(f)
col = {h: i for i, h in enumerate(hdr)}
names = {"grid": "Grid Size", "dur": "gpu__time_duration.sum", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
         "dr": "dram__bytes_read.sum", "dw": "dram__bytes_write.sum", "l2hit": "lts__t_sector_hit_rate.pct",
         "l2thr": "lts__throughput.avg.pct_of_peak_sustained_elapsed"}
groups = collections.OrderedDict()
for row in f:
    k = re.sub(r"\(.*", "", row[col["Kernel Name"]]).replace("void <unnamed>::", "")
    key = (k, row[col[names["grid"]]])
    g = groups.setdefault(key, [])
    g.append([float(row[col[names[x]]].replace(",", "")) for x in ("dur", "tensor", "dr", "dw", "l2hit", "l2thr")])
w(f"## ncu --set full capture of tensor-core launches of the support forward pass (raw page: `profiles/{tag}_ncu_gemm_raw.csv`)\n")
w("| kernel | grid | launches | duration us | tensor pipe active % | DRAM read MB | DRAM write KB | L2 hit % | L2 throughput % | what it is |")
w("|---|---|---|---|---|---|---|---|---|---|")
what = {"(7, 2, 4)": "fused attention, 7 query (key) tiles x 2 heads x 4 utterances", "(1, 2, 4)": "fused attention, encoder (128 phonemes)",
        "(84, 1, 1)": "QKV projection 3456x768x256", "(49, 1, 8)": "attention scores 864x864x128 x 8 (b,h)", "(14, 1, 8)": "P.V 864x128x864 x 8",
        "(108, 1, 1)": "out-proj 3456x256x256 / conv k=1 3456x256x1024", "(32, 1, 4)": "conv k=9 864x1024x(9x256) x 4", "(4, 1, 4)": "encoder conv k=1 (tail of the query encoder branch)"}
for (k, grid), rows in groups.items():
    m = [sum(c) / len(rows) for c in zip(*rows)]
    w(f"| `{k}` | {grid} | {len(rows)} | {m[0]:.1f} | {m[1]:.1f} | {m[2]:.2f} | {m[3]:.1f} | {m[4]:.0f} | {m[5]:.0f} | {what.get(grid, '')} |")
open(P + f"{tag}_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])

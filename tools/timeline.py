"""Where does the wall time of one outer step go?  CUPTI (kineto) start / end stamps of every kernel of one CUDA-graph replay of
the task step + optimizer step: per kernel family the time during which it is the ONLY kernel running ('exclusive': it is on the
critical path by construction), the time it overlaps others, and the idle gaps (no kernel running at all).

    python tools/timeline.py [--workload config2]
"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
from meta_tts_b200 import synthetic as SYN  # noqa: E402
from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import copy  # noqa: E402

if "--workload" in sys.argv:
    B.set_workload(sys.argv[sys.argv.index("--workload") + 1])
algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
algo["adapt"]["train"]["steps"] = B.K_INNER
algo["adapt"]["test"]["steps"] = B.K_INNER
sysm = MetaSystem(None, DEFAULT_MODEL_CONFIG, copy.deepcopy(DEFAULT_TRAIN_CONFIG), algo, n_speaker=16, device="cuda:0", split=3, dropout=True,
                  seed=0, second_order=not B.FIRST_ORDER)
sysm.graph_min_hits = 1
sysm.load_state_dict({k: v.detach() for k, v in SYN.init_state_dict(DEFAULT_MODEL_CONFIG, n_speaker=16, seed=0).items()})
t = SYN.synth_task(task=0, shots=B.SHOTS, queries=B.QUERIES, L=B.L_PHON, T=B.T_MEL)
sysm.training_step([([t[0]], [t[1]])], 0)
sysm.optimizer_step()
graph = next(iter(sysm._graphs.values()))[2]
for _ in range(3):
    graph.replay()
    sysm.optimizer_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    graph.replay()
    sysm.optimizer_step()
    torch.cuda.synchronize()
evs = []
for ev in prof.events():
    if ev.device_type.name != "CUDA" or "memcpy" in ev.name.lower() or "memset" in ev.name.lower():
        continue
    evs.append((ev.time_range.start, ev.time_range.end, ev.name))
evs.sort()
t0, t1 = evs[0][0], max(e[1] for e in evs)


def fam(n):
    n = re.sub(r"\(anonymous namespace\)::|void ", "", n)
    m = re.match(r"([A-Za-z_0-9:]+(<[^>]*>)?)", n)
    return m.group(1) if m else n[:40]


# sweep line over start / end points
pts = []
for i, (a, b, n) in enumerate(evs):
    pts.append((a, 1, i))
    pts.append((b, -1, i))
pts.sort()
active = set()
excl = collections.Counter()
shared = collections.Counter()
idle = 0.0
prev = pts[0][0]
for tpt, kind, i in pts:
    dt = tpt - prev
    if dt > 0:
        if not active:
            idle += dt
        elif len(active) == 1:
            excl[fam(evs[next(iter(active))][2])] += dt
        else:
            for j in active:
                shared[fam(evs[j][2])] += dt / len(active)
    prev = tpt
    if kind == 1:
        active.add(i)
    else:
        active.discard(i)
wall = t1 - t0
print(f"# one outer step: {len(evs)} kernels, wall {wall / 1e3:.3f} ms (under the profiler), no kernel running {idle / 1e3:.3f} ms ({100 * idle / wall:.1f}%)")
print("| kernel family | launches | exclusive ms (alone on the GPU) | shared ms (its share while overlapping) | % of wall |")
print("|---|---|---|---|---|")
cnt = collections.Counter(fam(e[2]) for e in evs)
for k in sorted(cnt, key=lambda k: -(excl[k] + shared[k])):
    if excl[k] + shared[k] < 0.01 * wall:
        continue
    print(f"| `{k}` | {cnt[k]} | {excl[k] / 1e3:.3f} | {shared[k] / 1e3:.3f} | {100 * (excl[k] + shared[k]) / wall:.1f} |")

# second table: exclusive time by (family, duration bucket) — which individual launches sit alone on the critical path
if "--detail" in sys.argv:
    excl_i = collections.Counter()
    active, prev = set(), pts[0][0]
    for tpt, kind, i in pts:
        if tpt > prev and len(active) == 1:
            excl_i[next(iter(active))] += tpt - prev
        prev = tpt
        (active.add if kind == 1 else active.discard)(i)
    groups = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for i, (a, b, n) in enumerate(evs):
        dur = b - a
        bucket = 2 ** max(0, int(dur).bit_length() - 1)          # power-of-two duration bucket (us)
        g = groups[(fam(n), bucket)]
        g[0] += 1
        g[1] += dur
        g[2] += excl_i[i]
    print("\n| kernel family | duration bucket us | launches | mean us | exclusive ms |")
    print("|---|---|---|---|---|")
    for (k, bkt), (c, d, e) in sorted(groups.items(), key=lambda kv: -kv[1][2])[:45]:
        print(f"| `{k}` | {bkt}-{2 * bkt} | {c} | {d / c:.1f} | {e / 1e3:.3f} |")

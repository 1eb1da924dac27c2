"""Timeline of one CUDA-graph replay of the C2 training step (torch.profiler / CUPTI): span, idle gaps, and for every
kernel name the time it runs alone vs overlapped.  Diagnostic only."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from msmp_pde_b200 import models_gnn2D, synth
from msmp_pde_b200.train_step import GraphedTrainStep

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 64))
pde, data, meta = synth.config_c2(B=B, nx=100, seed=0)
torch.manual_seed(0)
model = models_gnn2D.MP_PDE_Solver2DLEMLinGated(pde, 25, 128, 6, meta["eq_variables"]).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=True)
g = data.clone().to(dev)
step = GraphedTrainStep(model, opt, g, warmup=3)
for _ in range(5):
    step(g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(g)
        torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "trace.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
prof.export_chrome_trace(out)
tr = json.load(open(out))
ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("ph") == "X"]
ks.sort(key=lambda e: e["ts"])
# split into replays at the pack kernel (first msmp kernel of every step)
starts = [i for i, e in enumerate(ks) if "k_pack" in e["name"]]
lo = starts[-1]
while lo > 0 and ks[lo]["ts"] - (ks[lo - 1]["ts"] + ks[lo - 1]["dur"]) < 100:
    lo -= 1
grp = ks[lo:]
t0 = grp[0]["ts"]; t1 = max(e["ts"] + e["dur"] for e in grp)
print(f"replay: {len(grp)} kernels, span {(t1 - t0) / 1e3:.3f} ms, sum of durations {sum(e['dur'] for e in grp) / 1e3:.3f} ms, "
      f"streams {len(set(e['args'].get('stream') for e in grp))}")
# sweep: events
pts = []
for i, e in enumerate(grp):
    pts.append((e["ts"], 1, i)); pts.append((e["ts"] + e["dur"], -1, i))
pts.sort()
active = set(); last = t0; idle = 0.0
alone = collections.Counter(); shared = collections.Counter(); cnt = collections.Counter(); tot = collections.Counter()
short = lambda n: (n[:118] if "at::native" in n else n.split("(")[0].split("<")[0][-60:])
for e in grp:
    cnt[short(e["name"])] += 1; tot[short(e["name"])] += e["dur"]
for t, d, i in pts:
    dtm = t - last
    if dtm > 0:
        if not active:
            idle += dtm
        elif len(active) == 1:
            alone[short(grp[next(iter(active))]["name"])] += dtm
        else:
            for j in active:
                shared[short(grp[j]["name"])] += dtm / len(active)
    last = t
    if d == 1: active.add(i)
    else: active.discard(i)
print(f"idle (no kernel running): {idle / 1e3:.3f} ms")
print(f"{'kernel':60s} {'n':>5s} {'sum us':>9s} {'alone us':>9s} {'shared us':>9s}")
for n, _ in sorted(tot.items(), key=lambda kv: -(alone[kv[0]] + shared[kv[0]])):
    print(f"{n:60s} {cnt[n]:5d} {tot[n]:9.1f} {alone[n]:9.1f} {shared[n]:9.1f}")
# coarse phases: first/last occurrence of marker kernels
def first(name): return next((e["ts"] - t0 for e in grp if name in e["name"]), None)
def lastt(name): return next((e["ts"] + e["dur"] - t0 for e in reversed(grp) if name in e["name"]), None)
for nm in ("k_pack", "k_lem_fwd_tc", "k_edge_", "k_decoder_fwd", "k_decoder_bwd", "k_lem_bwd_tc", "multi_tensor", "adam"):
    print(f"  {nm:16s} first start {first(nm)} us, last end {lastt(nm)} us")

if os.environ.get("TAIL_US"):
    tail = float(os.environ["TAIL_US"])
    sid = {}
    print("--- kernels in the last %.0f us (start us, dur us, stream, name)" % tail)
    for e in grp:
        if e["ts"] + e["dur"] - t0 >= (t1 - t0) - tail:
            st = sid.setdefault(e["args"].get("stream"), len(sid))
            print(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f}  s{st}  {short(e['name'])}  grid={e['args'].get('grid')}")

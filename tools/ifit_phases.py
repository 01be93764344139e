"""ifit phase timers (cw_ifit.cu MARK() / FMARK()) for a run: where a level-step spends its time.
  python tools/ifit_phases.py [n] [d] [kind] [cluster]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from rag_cobweb_b200 import CobwebTorchTree, synth, _lib
n, d, kind = int(sys.argv[1]) if len(sys.argv) > 1 else 30000, int(sys.argv[2]) if len(sys.argv) > 2 else 768, sys.argv[3] if len(sys.argv) > 3 else "unit"
if len(sys.argv) > 4:
    _lib.check(_lib.load().cw_set_ifit_cluster(int(sys.argv[4])))
x = torch.from_numpy(synth.corpus(n, d, kind, 0)).cuda()
t = CobwebTorchTree((d,))
torch.cuda.synchronize(); t0 = time.time()
t.ifit_batch(x, tag_sentences=True)
torch.cuda.synchronize(); dt = time.time() - t0
c = t.store.counters()
ph = t.store.ifit_phase_cycles()
tot = sum(ph.values())
print(f"{n}x{d} {kind}: {n / dt:.0f} inserts/s, {c['levels'] / n:.2f} levels/insert, {dt / c['levels'] * 1e6:.2f} us/level-step, rows/insert {c['rows'] / n:.1f}")
if tot == 0:
    print("(phase timers not compiled in: CW_IFIT_FINE_TIMERS=1 python rag-cobweb_b200/build.py --force)")
print("phase share:", {k: round(v / max(tot, 1), 3) for k, v in ph.items()}, "cycles/level", {k: int(v / max(c['levels'], 1)) for k, v in ph.items()})
base = 16 + 4 * _lib.MAX_CHILDREN + 1 + 3
w = t.store.scratch[base:base + 48].cpu().numpy().view(np.int64)
names = {10: "phase B .. next entry", 11: "current node row load", 12: "slice work (this team: the cached rows)", 13: "wait for the P' slices", 14: "job setup", 15: "child row load",
         17: "job arithmetic", 18: "team reduce", 19: "sends + further rounds", 20: "exchange A", 21: "-", 22: "-",
         23: "decision A + best1 child list (waiting)"}
if w[10:24].any():
    print("fine (cycles/level, first thread of the team with job 0 in the lead CTA):", {names[k]: int(w[k] / max(c['levels'], 1)) for k in names})

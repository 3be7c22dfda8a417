"""Per-kernel timing at large n (HBM-resident).  Usage: python scratch/scale_bench.py [n]"""
import sys, json, time, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mocat_b200 import _lib, engine, models
from mocat_b200._lib import ptr, stream
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
PEAK = 6529.7
L = _lib.get(); ctx = L.ctx()
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
res = {}
def report(name, ms, bytes_per_particle, note=""):
    gbs = bytes_per_particle * n / (ms * 1e-3) / 1e9
    res[name] = dict(ms=ms, algo_bytes_per_particle=bytes_per_particle, gbs=gbs, frac=gbs / PEAK, note=note)
    print("%-34s %9.3f ms  %7.1f GB/s  %5.1f%% of %.0f   %s" % (name, ms, gbs, 100 * gbs / PEAK, PEAK, note), flush=True)

dev = torch.device("cuda")
lw = torch.randn(n, device=dev) * 2
lik = torch.rand(n, device=dev) * 20
out6 = torch.empty(6, dtype=torch.float64, device=dev)
report("lse_ess", timeit(lambda: L.call("mb_lse_ess", ctx, ptr(lw), None, 0.0, n, ptr(out6), stream())), 4)
report("lse_ess_tempered", timeit(lambda: L.call("mb_lse_ess", ctx, ptr(lw), ptr(lik), 0.3, n, ptr(out6), stream())), 8)
ctl = engine.ControlBlock()
o = out6.cpu().numpy()
L.call("mb_lse_ess", ctx, ptr(lw), None, 0.0, n, ptr(out6), stream()); o = out6.cpu().numpy()
rec = np.zeros(1, dtype=_lib.CONTROL_DTYPE)[0]; rec["wmax"], rec["s1"], rec["s2"], rec["resample"] = o[0], o[1], o[2], 1
ctl.write(rec)
cdf = torch.empty(n, dtype=torch.float64, device=dev)
report("scan_cdf (lw -> fp64 cdf)", timeit(lambda: L.call("mb_cumsum_lw", ctx, ptr(lw), n, ptr(ctl.t), 1, ptr(cdf), stream())), 12, "4 read + 8 write")
anc = torch.empty(n, dtype=torch.int32, device=dev)
report("ancestors systematic", timeit(lambda: L.call("mb_ancestors", ctx, ptr(cdf), n, 0, None, 1, 1, 0, ptr(anc), n, None, stream())), 12, "8 read + 4 write")
report("ancestors multinomial iid-u", timeit(lambda: L.call("mb_ancestors", ctx, ptr(cdf), n, 1, None, 1, 1, 0, ptr(anc), n, None, stream()), reps=2, warm=1), 12, "legacy: unsorted u, random binary search")
B = int(L.dll.mb_strata_count(n)); hist = torch.zeros(B, dtype=torch.int32, device=dev); offs = torch.zeros(B + 1, dtype=torch.int32, device=dev)
def strat():
    L.call("mb_strata_hist", ctx, n, 0, B, 1, 1, None, ptr(hist), 1, stream())
    L.call("mb_ancestors_sorted", ctx, ptr(cdf), n, None, 1, ptr(hist), ptr(offs), B, 1, 1, 0, n, ptr(anc), n, None, stream())
report("ancestors multinomial stratified", timeit(strat), 12, "hist + scan + sorted search (B=%d)" % B)
report("ancestors systematic (sorted kernel)", timeit(lambda: L.call("mb_ancestors_sorted", ctx, ptr(cdf), n, None, 0, ptr(hist), ptr(offs), B, 1, 1, 0, n, ptr(anc), n, None, stream())), 12, "8 read + 4 write")
L.call("mb_ancestors", ctx, ptr(cdf), n, 0, None, 1, 1, 0, ptr(anc), n, None, stream())
d = 5
ld = (n + 31) // 32 * 32
src = torch.randn((d, ld), device=dev); 
report("gather_state d=5 (sorted anc)", timeit(lambda: engine.gather_state(anc, src, n)), 4 + 8 * d, "anc + 2*4*d")
del src
torch.cuda.empty_cache()
# SMC move
tgt = models.make_target(_lib.LIK_RASTRIGIN, 5, prior_std=3.0, a=1.0)
eng = engine.SMCEngine(tgt, models.make_move(_lib.MOVE_MALA, 0.1), models.make_temper(max_iter=1 << 30), n, 1)
eng.use_graphs = False
eng.startup()
def step_move():
    src_, dst_ = eng.xbuf[eng.cur], eng.xbuf[eng.cur ^ 1]
    L.call("mb_smc_move", ctx, C.byref(eng.target), C.byref(eng.move), ptr(src_), ptr(dst_), eng.ld, n, ptr(eng.anc), ptr(eng.lw), ptr(eng.up), ptr(eng.lik), ptr(eng.alpha), 1, 0, ptr(eng.ctl.t), None, stream())
report("smc_move Rastrigin d=5 MALA", timeit(step_move), 64, "contract 64 B (actual 4d r + 4d+12 w = 52)")
report("temper_adapt (search+update)", timeit(lambda: eng._temper(True)), 8 * 6 + 12, "8 B x ~6 evals + 12")
c = eng.ctl.read(); print("   search iters", c['search_iters'], "beta", c['beta'])
ms_steps = []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.update(); e1.record(); torch.cuda.synchronize(); ms_steps.append(e0.elapsed_time(e1))
print("   full SMC steps ms:", ["%.2f" % m for m in ms_steps], " => %.3g particle-steps/s" % (n / (np.median(ms_steps) * 1e-3)))
res["smc_full_step_ms"] = float(np.median(ms_steps))
del eng; torch.cuda.empty_cache()
# Lorenz-96 PF, d = 40
d = 40
s = models.make_lorenz96(dim=d)
pf = engine.PFEngine(s, n, 3, ess_threshold=2.0, resampling=_lib.RESAMPLE_SYSTEMATIC)
y = torch.randn(d, device=dev) + 2
pf.init(y)
def pf_only():
    src_, dst_ = pf.xbuf[pf.cur], pf.xbuf[pf.cur ^ 1]
    L.call("mb_pf_step", ctx, C.byref(pf.ssm), ptr(src_), ptr(dst_), pf.ld, n, n, ptr(pf.anc), ptr(y), ptr(pf.lw), 3, 1, 0, 2.0, ptr(pf.ctl.t), ptr(pf.ctl.hist), None, None, stream())
rec = pf.ctl.read(); rec['resample'] = 0; pf.ctl.write(rec)
report("pf_step L96 d=40 (no resample)", timeit(pf_only, reps=3, warm=1), 8 * d + 8, "x r/w + lw r/w")
ms_steps = []
for _ in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pf.step(y); e1.record(); torch.cuda.synchronize(); ms_steps.append(e0.elapsed_time(e1))
msf = float(np.median(ms_steps))
report("PF full step L96 (resample every step)", msf, 336, "contract 2*4*d+16")
res["pf_particle_steps_per_s"] = n / (msf * 1e-3)
del pf; torch.cuda.empty_cache()
# ABC
data = np.sort(np.random.default_rng(0).standard_normal(8) * 2 + 3)
ab = engine.ABCEngine(models.make_gk(data), n, 5, max_iter=1000)
ab.startup()
ms_steps = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ab.update(); e1.record(); torch.cuda.synchronize(); ms_steps.append(e0.elapsed_time(e1))
report("ABC full step g-and-k m=8", float(np.median(ms_steps)), 60, "contract 60 B move; + quantile/colstats/resample")
json.dump(res, open("/root/repo/gpurun_out/scale_%g.json" % n, "w"), indent=1)

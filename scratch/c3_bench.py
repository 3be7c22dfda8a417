"""Per-kernel timing of the C3 path at n particles (default 1e8): python scratch/c3_bench.py [n] [r_std]"""
import sys, json, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mocat_b200 import _lib, engine, models
from mocat_b200._lib import ptr, stream
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
r_std = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
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
def report(name, ms, bpp, note=""):
    gbs = bpp * n / (ms * 1e-3) / 1e9
    print("%-44s %9.3f ms  %7.1f GB/s  %5.1f%% of %.0f   %s" % (name, ms, gbs, 100 * gbs / PEAK, PEAK, note), flush=True)
d = 40
s = models.make_lorenz96(dim=d, r_std=r_std)
pf = engine.PFEngine(s, n, 3, ess_threshold=2.0, resampling=_lib.RESAMPLE_SYSTEMATIC)
from oracle import models as om
_, ys = om.Lorenz96SSM(dim=d, r_std=r_std).simulate(30, np.random.default_rng(0), spinup=1000)
yd = torch.as_tensor(ys.astype(np.float32), device="cuda")
pf.init(yd[0])
for t in range(1, 4): pf.step(yd[t])
c = pf.ctl.read(); print("ess after 3 steps", c['ess'], "log_z", c['log_z'])
st = stream()
def res_a(): L.call("mb_rs_tile_sums", ctx, ptr(pf.rs_ws), ptr(pf.lw), n, n, 1, ptr(pf.ctl.t), 1, st)
def res_b(): L.call("mb_rs_ancestors", ctx, ptr(pf.rs_ws), ptr(pf.lw), n, n, 1, ptr(pf.ctl.t), 1, -1, None, None, ptr(pf.anc), st)
report("rs_tile_sums (degenerate weights)", timeit(res_a), 4)
report("rs_ancestors + heavy (degenerate weights)", timeit(res_b), 8)
hc = int(pf.rs_ws[1].item() & 0xffffffff); print("   heavy tiles:", hc)
def step_only():
    pf.t += 1; pf._step_kernel(yd[4], st); pf.cur ^= 1; pf.t -= 1      # same buffers every time
rec = pf.ctl.read(); rec['resample'] = 1; pf.ctl.write(rec)
def step_res():
    r = pf.ctl.read(); 
    pf._step_kernel(yd[4], st); pf.cur ^= 1
report("pf_l96 step, gather by degenerate ancestors", timeit(lambda: (pf._step_kernel(yd[4], st), setattr(pf, 'cur', pf.cur ^ 1))), 8 * d + 12)
rec = pf.ctl.read(); rec['resample'] = 0; pf.ctl.write(rec)
def norestep():
    rec2 = None
    pf._step_kernel(yd[4], st); pf.cur ^= 1
    # the step sets ctl.resample again (threshold 2.0): clear it on the stream-ordered host path
# no-resample variant: threshold 0 so the kernel keeps resample = 0
pf.ess_threshold = 0.0
pf._step_kernel(yd[4], st); pf.cur ^= 1
report("pf_l96 step, no resampling (streams x)", timeit(lambda: (pf._step_kernel(yd[4], st), setattr(pf, 'cur', pf.cur ^ 1))), 8 * d + 8)
pf.ess_threshold = 2.0
# flat weights: lw = small noise -> every particle ~1 offspring
pf._lw_full[:n] = torch.randn(n, device="cuda") * 0.3
L.call("mb_lse_ess", ctx, ptr(pf.lw), None, 0.0, n, ptr(torch.empty(6, dtype=torch.float64, device="cuda")), st)
out6 = torch.empty(6, dtype=torch.float64, device="cuda"); L.call("mb_lse_ess", ctx, ptr(pf.lw), None, 0.0, n, ptr(out6), st)
o = out6.cpu().numpy(); rec = pf.ctl.read(); rec['wmax'], rec['s1'], rec['s2'], rec['resample'] = o[0], o[1], o[2], 1; pf.ctl.write(rec)
report("rs_tile_sums (flat weights)", timeit(res_a), 4)
report("rs_ancestors + heavy (flat weights)", timeit(res_b), 8)
hc = int(pf.rs_ws[1].item() & 0xffffffff); print("   heavy tiles:", hc)
lwsave = pf._lw_full.clone()
def step_flat():
    pf._lw_full.copy_(lwsave)
report("lw restore copy (subtract)", timeit(lambda: pf._lw_full.copy_(lwsave)), 8)
def sf():
    r = pf._step_kernel(yd[4], st); pf.cur ^= 1
rec = pf.ctl.read(); rec['resample'] = 1; pf.ctl.write(rec)
# a step with resample=1 rewrites lw and ctl (resample stays 1 with threshold 2.0); ancestors stay the flat ones
report("pf_l96 step, gather by flat ancestors", timeit(sf), 8 * d + 12)
# gather kernels
src, dst = pf.xbuf[pf.cur], pf.xbuf[pf.cur ^ 1]
report("gather_rows staged (flat anc)", timeit(lambda: L.call("mb_gather_rows", ctx, ptr(pf.anc), n, d, ptr(src), n, ptr(dst), 1, st)), 8 * d + 4)
report("gather_rows direct (flat anc)", timeit(lambda: L.call("mb_gather_rows", ctx, ptr(pf.anc), n, d, ptr(src), n, ptr(dst), 0, st)), 8 * d + 4)
report("weighted_moments_rows (flat)", timeit(lambda: pf.moments()), 4 * d + 4)
# full steps
for thr, nm in ((2.0, "every step"),):
    ms = []
    pf.ess_threshold = thr
    for t in range(5, 15):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pf.step(yd[t]); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    report("PF full step L96 (resample %s)" % nm, float(np.median(ms[2:])), 336, "contract 2*4*d+16; steps: " + " ".join("%.2f" % m for m in ms))
    print("   ess", pf.ctl.read()['ess'])

import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mocat_b200 import _lib as l, engine as e, models as m
from oracle import models as omodels
lib = l.get()
def P(*a):
    torch.cuda.synchronize(); print(*a, flush=True)
d, world, nl, seed = 40, 4, 32 * 320, 9
n = world * nl
s = m.make_lorenz96(dim=d)
_, y = omodels.Lorenz96SSM(dim=d).simulate(2, np.random.default_rng(0), spinup=100)
yd = torch.as_tensor(y.astype(np.float32), device="cuda")
for collapse in (False, True):
    eng = e.PFEngine(s, n, seed, ess_threshold=2.0, resampling=l.RESAMPLE_SYSTEMATIC)
    eng.init(yd[0])
    if not collapse:
        eng._lw_full.mul_(0.02)
        o = e.lse_ess(eng.lw).cpu().numpy()
        c = eng.ctl.read(); c['wmax'], c['s1'], c['s2'] = o[0], o[1], o[2]; eng.ctl.write(c)
    x0, lw0, ctl0 = eng.x.clone(), eng._lw_full.clone(), eng.ctl.t.clone()
    P("ref step", collapse)
    eng.step(yd[1])
    P("ref done")
    x_ref, anc_ref, lw_ref = eng.x.clone(), eng.anc.clone(), eng._lw_full.clone()
    tiles = nl // 32
    stride = d + 8
    anc = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    imp = torch.zeros((n, stride), dtype=torch.float32, device="cuda")
    wss = [torch.zeros((int(lib.dll.mb_rs_workspace_bytes(nl)) + 7) // 8, dtype=torch.int64, device="cuda") for _ in range(world)]
    ctls = []
    for r in range(world):
        ctl = e.ControlBlock(); ctl.t.copy_(ctl0); ctls.append(ctl)
        lib.call("mb_rs_tile_sums", lib.ctx(), l.ptr(wss[r]), l.ptr(lw0[r * nl:]), nl, n, 1, l.ptr(ctl.t), 0, l.stream())
    P("tile sums done")
    totals = torch.stack([ws[0] for ws in wss]).contiguous()
    shards = []
    for r in range(world):
        sh = l.Shard()
        sh.rank, sh.world, sh.n_local, sh.n_total = r, world, nl, n
        for q in range(world):
            sh.x_peers[q] = x0[q * tiles:].data_ptr()
            sh.anc_peers[q] = anc[q * nl:].data_ptr()
            sh.lw_peers[q] = lw0[q * nl:].data_ptr()
            sh.ws_peers[q] = wss[q].data_ptr()
            sh.import_peers[q] = imp[q * nl:].data_ptr()
        sh.import_stride, sh.state_dim = stride, d
        shards.append(sh)
    for r in range(world):
        lib.call("mb_rs_ancestors", lib.ctx(), l.ptr(wss[r]), l.ptr(lw0[r * nl:]), nl, n, 1, l.ptr(ctls[r].t), 0, -1,
                 l.ptr(totals), C.byref(shards[r]), l.ptr(anc[r * nl:]), l.stream())
        P("pass B", r)
    for r in range(world):
        lib.call("mb_rs_heavy", lib.ctx(), l.ptr(wss[r]), l.ptr(lw0[r * nl:]), nl, n, 1, l.ptr(ctls[r].t), 0, -1,
                 l.ptr(totals), C.byref(shards[r]), l.ptr(anc[r * nl:]), l.stream())
        P("pass C", r, "heavy counts", [int(ws[1].item() & 0xffffffff) for ws in wss])
    print("anc equal", torch.equal(anc, anc_ref), int((anc != anc_ref).sum()), flush=True)
    x_out = torch.zeros_like(x0)
    lw = lw0.clone()
    for r in range(world):
        lib.call("mb_pf_l96_step", lib.ctx(), C.byref(s), l.ptr(x0[r * tiles:]), l.ptr(x_out[r * tiles:]), nl, n,
                 l.ptr(anc[r * nl:]), l.ptr(yd[1]), l.ptr(lw[r * nl:]), seed, 1, r * nl, 2.0, l.ptr(ctls[r].t), None,
                 C.byref(shards[r]), None, l.stream())
        P("step", r)
    print("lw equal", torch.equal(lw[:n], lw_ref[:n]), "x equal", torch.equal(x_out, x_ref), flush=True)

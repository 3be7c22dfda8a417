"""torchrun worker: sharded SMC / PF over WORLD_SIZE GPUs must reproduce the single-GPU population.
Launched by tests/test_gpu_multi.py:  torchrun --nproc-per-node N tests/mp_sharded_worker.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocat_b200 import _lib, engine, models, parallel  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    sc = parallel.ShardContext()
    ok = True

    # ---- mailbox allgather
    src = torch.tensor([rank + 0.5, 10.0 * rank, -1.0], dtype=torch.float64, device="cuda")
    dst = torch.zeros(3 * world, dtype=torch.float64, device="cuda")
    for rep in range(5):
        sc.allgather(src + rep, 3, dst)
        exp = np.concatenate([[r + 0.5 + rep, 10.0 * r + rep, -1.0 + rep] for r in range(world)])
        ok &= bool(np.array_equal(dst.cpu().numpy(), exp))
    print(f"[rank {rank}] allgather ok={ok}", flush=True)

    # ---- tempered SMC, systematic + multinomial: step-by-step against the single-GPU engine.
    # Reductions are partitioned differently (fp32 partial sums per block), so temperatures agree to ~1e-8
    # relative, not bitwise; particles are bit-identical until the first resampling, where >= 99 % of the
    # ancestors must coincide (exact CDF; the scale differs by ~1e-9), afterwards the runs are compared
    # statistically.
    n_local, d, seed = 40_000, 5, 9

    def gather(t):
        t = t.contiguous()
        g = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        return torch.cat(g).cpu().numpy()

    for resampling in (_lib.RESAMPLE_SYSTEMATIC, _lib.RESAMPLE_MULTINOMIAL):
        tgt = models.make_target(_lib.LIK_RASTRIGIN, d, prior_std=3.0, a=1.0)
        mk = lambda: (models.make_move(_lib.MOVE_MALA, 0.1), models.make_temper(max_iter=25))
        eng = parallel.ShardedSMCEngine(sc, tgt, *mk(), n_local, seed, resampling=resampling)
        ref = None
        if rank == 0:
            ref = engine.SMCEngine(tgt, *mk(), n_local * world, seed, resampling=resampling)
            ref.use_graphs = False
            ref.startup()
        eng.startup()
        seen_resample = False
        for it in range(25):
            eng.update()
            xg, ag = gather(eng.values()), gather(eng.anc)
            c = eng.ctl.read()
            if rank == 0:
                ref.update()
                cr = ref.ctl.read()
                ok &= c['iter'] == cr['iter'] and c['resampled'] == cr['resampled']
                if not seen_resample and not c['resampled']:
                    ok_b = abs(c['beta'] - cr['beta']) < 1e-6 * cr['beta'] and abs(c['ess'] - cr['ess']) < 1e-6 * cr['ess']
                    # the temperatures agree to ~1e-9 (different partition of the fp32 block sums); when that flips the
                    # fp32 rounding of beta the particles move by an ulp, so: identical up to rounding noise
                    xr = ref.values().cpu().numpy()
                    same_x = float(np.mean(xg == xr))
                    ok_x = same_x > 0.98 and float(np.max(np.abs(xg - xr))) < 1e-3
                    if not (ok_b and ok_x):
                        print(f"[smc mode {resampling}] iter {it + 1}: beta {c['beta']!r} vs {cr['beta']!r} ess {c['ess']!r} vs "
                              f"{cr['ess']!r} identical x {same_x:.5f} max diff {np.max(np.abs(xg - xr)):.3g}", flush=True)
                    ok &= ok_b and ok_x
                elif not seen_resample:
                    same_anc = float(np.mean(ag == ref.anc.cpu().numpy()))
                    print(f"[smc mode {resampling}] first resampling at iter {it + 1}: ancestors identical {same_anc:.5f}",
                          flush=True)
                    ok &= same_anc > 0.99
                    seen_resample = True
        c = eng.ctl.read()
        if rank == 0:
            cr = ref.ctl.read()
            print(f"[smc mode {resampling}] final beta {c['beta']:.6f} vs {cr['beta']:.6f}  log_z {c['log_z']:.4f} vs "
                  f"{cr['log_z']:.4f} ok={ok}", flush=True)
            ok &= seen_resample and abs(c['beta'] - cr['beta']) < 2e-2 * cr['beta'] and abs(c['log_z'] - cr['log_z']) < 0.05
        dist.barrier()

    # ---- bootstrap PF on Lorenz-96 d=8, resampling every step
    from oracle import models as omodels
    d = 8
    _, y = omodels.Lorenz96SSM(dim=d).simulate(8, np.random.default_rng(0), spinup=200)
    yd = torch.as_tensor(y.astype(np.float32), device="cuda")
    s = models.make_lorenz96(dim=d)
    pf = parallel.ShardedPFEngine(sc, s, 32_000, 5, ess_threshold=2.0, resampling=_lib.RESAMPLE_SYSTEMATIC)
    pf.init(yd[0])
    for t in range(1, len(y)):
        pf.step(yd[t])
    c = pf.ctl.read()
    xs = pf.values().contiguous()
    gathered = [torch.empty_like(xs) for _ in range(world)]
    dist.all_gather(gathered, xs)
    if rank == 0:
        ref = engine.PFEngine(s, 32_000 * world, 5, ess_threshold=2.0, resampling=_lib.RESAMPLE_SYSTEMATIC)
        ref.init(yd[0])
        for t in range(1, len(y)):
            ref.step(yd[t])
        cr = ref.ctl.read()
        same = float(np.mean(np.all(torch.cat(gathered).cpu().numpy() == ref.values().cpu().numpy(), axis=1)))
        print(f"[pf] log_z {c['log_z']:.6f} vs {cr['log_z']:.6f}; ess {c['ess']:.3f} vs {cr['ess']:.3f}; "
              f"identical particles {same:.5f}", flush=True)
        ok &= abs(c['log_z'] - cr['log_z']) < 1e-6 * abs(cr['log_z']) + 1e-6 and same > 0.995
    dist.barrier()

    # ---- SMC-ABC on the g-and-k model (config C5's shape, smaller): global quantile threshold / ESS / acceptance of the
    # sharded population against the single-GPU engine.  The first threshold is bit-identical (same Philox streams, exact
    # radix select over the ranks' histograms); later ones agree to rounding (the column variances are summed in a
    # different order, which can move a proposal by an ulp).
    zz = np.random.default_rng(0).standard_normal(8)
    data = np.sort(3.0 + 1.0 * (1 + 0.8 * np.tanh(2.0 * zz / 2)) * zz * (1 + zz * zz) ** 0.5)
    n_a = 40_000
    abc = parallel.ShardedABCEngine(sc, models.make_gk(data), n_a, 21, max_iter=12)
    abc.startup()
    ref_abc = None
    if rank == 0:
        ref_abc = engine.ABCEngine(models.make_gk(data), n_a * world, 21, max_iter=12)
        ref_abc.startup()
        c, cr = abc.ctl.read(), ref_abc.ctl.read()
        ok_a = c['beta'] == cr['beta'] and c['ess'] == cr['ess']
        print(f"[abc] startup threshold {c['beta']!r} vs {cr['beta']!r}, ess {c['ess']} vs {cr['ess']} ok={ok_a}", flush=True)
        ok &= bool(ok_a)
    for it in range(8):
        abc.update()
        if rank == 0:
            ref_abc.update()
    c = abc.ctl.read()
    if rank == 0:
        cr = ref_abc.ctl.read()
        ok_a = (c['iter'] == cr['iter'] and abs(c['beta'] - cr['beta']) < 2e-3 * abs(cr['beta']) and
                abs(c['ess'] - cr['ess']) < 5e-3 * cr['ess'] and abs(c['alpha_mean'] - cr['alpha_mean']) < 5e-3)
        print(f"[abc] iter {c['iter']}: threshold {c['beta']:.6f} vs {cr['beta']:.6f}, ess {c['ess']:.0f} vs {cr['ess']:.0f}, "
              f"alpha {c['alpha_mean']:.4f} vs {cr['alpha_mean']:.4f} ok={ok_a}", flush=True)
        ok &= bool(ok_a)
    dist.barrier()

    # ---- SVGD (config C4's shape, smaller): the ensemble sharded by rows over the ranks reproduces the replicated run.
    # phi rows, adagrad and the logistic gradient are row-wise identical; the median bandwidth adds up integer counters.
    import mocat_b200 as mocat
    from mocat_b200 import kernels
    rng = np.random.default_rng(3)
    n_s, d_s, N_s = 4096, 50, 512
    A = rng.standard_normal((N_s, d_s)).astype(np.float32)
    lab = (rng.random(N_s) < 1.0 / (1.0 + np.exp(-(A @ rng.standard_normal(d_s))))).astype(np.float32)
    sc_s = mocat.scenarios.LogisticRegression(A, lab)

    class SVGDMedian(mocat.SVGD):
        def adapt(self, st, extra):
            extra.parameters.kernel_params.bandwidth = kernels.median_bandwidth_update(st.value)
            return st, extra
    out_sh = mocat.run(sc_s, SVGDMedian(max_iter=6, stepsize=0.05, keep_history=False), n=n_s, random_key=4)
    os.environ["MOCAT_B200_SVGD_SHARD"] = "0"
    out_rep = mocat.run(sc_s, SVGDMedian(max_iter=6, stepsize=0.05, keep_history=False, sharded=False), n=n_s, random_key=4)
    os.environ["MOCAT_B200_SVGD_SHARD"] = "1"
    dv = float(np.max(np.abs(out_sh.value - out_rep.value)))
    if rank == 0:
        print(f"[svgd] sharded vs replicated: max |dx| {dv:.3g}, bandwidth {out_sh.bandwidth:.7f} vs {out_rep.bandwidth:.7f}",
              flush=True)
    ok &= dv < 1e-5 and abs(out_sh.bandwidth - out_rep.bandwidth) < 1e-6 * out_rep.bandwidth
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if flag.item() == 1 else "SHARDED_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()

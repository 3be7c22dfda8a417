#!/usr/bin/env python
"""bench.py -- particle-steps/s of the tempered-SMC population step (BASELINE.json configs[1]:
adaptive likelihood tempering on Rastrigin d=5, n=1e6 particles per GPU, MALA moves).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference (oracle/)

One "step" = one SMCSampler.update (transport/smc.py:73-99): resample-if-needed (fp64 CDF scan, ancestor
search, gather fused into the move), MALA move with potential/gradient evaluation, adaptive temperature
search (regula falsi on device), weight update, ESS/log-evidence.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec"
D = 5
ALGO_BYTES_MOVE = 64.0          # x(5)+w+l+U_prior read + write
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the two step kernels at n = 1e6, one `ncu --set full`
# capture of this command (profiles/ncu_step_r1d.md); scaled linearly for other n.  The state is L2 resident, the
# writes of both kernels stay in L2, so the traffic is BELOW the algorithmic bytes (no wasted re-reads).
NCU_DRAM_BYTES_PER_PARTICLE = {"temper_adapt_kernel<resident>": 8.10, "smc_move_kernel<Rastrigin,5,MALA>": 24.27}
WORKLOAD = "C2 tempered SMC, Rastrigin d=5 a=1, prior N(0,3^2 I), MALA eps=0.1 (1 leapfrog), adaptive " \
           "tempering retain 0.9 / resample 0.5, multinomial resampling"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n", type=int, default=1_000_000, help="particles per GPU")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--seed", type=int, default=0)
    return p.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_arm(n, steps, warmup, seed, workers=None):
    """times oracle.parallel.ParallelTemperedSMC (NumPy restatement of transport/smc.py, all host cores)"""
    from oracle import models as om, parallel
    workers = workers or os.cpu_count() or 1
    s = parallel.ParallelTemperedSMC(om.IsoGaussianPrior(D, 0.0, 3.0), om.Rastrigin(D, 1.0), n, seed, move='mala',
                                     stepsize=0.1, max_iter=10000, workers=workers)
    st = s.startup()
    for _ in range(warmup):
        st = s.update(st) if not s.terminated(st) else s.startup()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        if s.terminated(st):
            st = s.startup()
        st = s.update(st)
        done += 1
    dt = time.perf_counter() - t0
    s.close()
    return n * done / dt, dt / done, workers


def reference_main(a, rank, world):
    if rank != 0:
        return
    n_sample = min(a.n, 1_000_000)
    steps = max(1, min(a.steps, 8))
    val, sec, workers = cpu_arm(n_sample, steps, min(a.warmup, 1), a.seed)
    sample = f"n={n_sample} particles x {steps} population steps of the same workload, NumPy restatement of the " \
             f"reference algorithm (mocat's JAX path cannot run: jax absent), {workers} worker processes"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "particle-steps/s", "n_gpus": a.gpus,
            "steps": steps, "warmup": min(a.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_per_step": n_sample, "dim": D},
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        return reference_main(a, rank, world)

    import torch
    import torch.distributed as dist
    import mocat_b200 as mocat
    from mocat_b200 import _lib, engine, models

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = a.n
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    tgt = models.make_target(_lib.LIK_RASTRIGIN, D, prior_std=3.0, a=1.0)
    if world > 1:
        # ONE population of world*n particles sharded over the GPUs (mocat_b200/parallel.py): LSE/ESS triples and
        # weight totals exchanged through peer-mapped mailboxes, ancestors gathered over NVLink peer reads
        from mocat_b200 import parallel
        sc = parallel.shard_context()
        eng = parallel.ShardedSMCEngine(sc, tgt, models.make_move(_lib.MOVE_MALA, 0.1),
                                        models.make_temper(max_iter=1 << 30), n, a.seed,
                                        resampling=_lib.RESAMPLE_MULTINOMIAL)
    else:
        eng = engine.SMCEngine(tgt, models.make_move(_lib.MOVE_MALA, 0.1), models.make_temper(max_iter=1 << 30), n,
                               a.seed, resampling=_lib.RESAMPLE_MULTINOMIAL)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def one_step(inner):
        """one population step bracketed by events; inner=True records per-kernel-group events (plain
        launches), inner=False replays the captured CUDA graph of the step (the production path)."""
        if eng.ctl.read()['done']:
            eng.startup()                                               # population reached beta = 1: start over (untimed)
            eng.update()                                                # first update after a restart is never graph-replayed
        flush.zero_()                                                   # L2 flush between timed iterations
        if inner:
            marks = [ev() for _ in range(4)]
            eng.update(events=marks)
            return marks
        e0, e1 = ev(), ev()
        e0.record()
        eng.update()
        e1.record()
        return [e0, e1]

    eng.startup()
    eng.update()
    sync()
    clocks = ClockSampler(local)
    clocks.start()
    t_w = time.perf_counter()
    nw = 0
    while True:                                                         # >= 1 s of warm-up so clocks settle
        for _ in range(max(a.warmup, 3) if nw == 0 else 100):
            one_step(False)
            nw += 1
        # the step count must be IDENTICAL on every rank (each step is a cross-GPU exchange): decide collectively
        el = torch.tensor([time.perf_counter() - t_w], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MIN)
        if float(el.item()) >= 1.0:
            break
    sync()
    all_marks = []
    t_wall0 = time.perf_counter()
    for _ in range(a.steps):
        all_marks.append(one_step(False))
    sync()
    t_wall = time.perf_counter() - t_wall0
    # second region: same steps with plain launches and per-kernel-group events (roofline of the kernels)
    inner_marks = [one_step(True) for _ in range(min(a.steps, 50))]
    sync()
    clk = clocks.stop()
    it_now = int(eng.ctl.read()['iter'])
    hist = eng.ctl.read_hist(it_now + 1)
    mean_search = float(hist['search_iters'][1:].mean()) if it_now >= 1 else 0.0
    # device time: per-step event pairs (flush excluded), summed
    tot = sum(m[0].elapsed_time(m[1]) for m in all_marks)              # ms
    ni = len(inner_marks)
    t_resample = sum(m[0].elapsed_time(m[1]) for m in inner_marks) / ni
    t_move = sum(m[1].elapsed_time(m[2]) for m in inner_marks) / ni
    t_temper = sum(m[2].elapsed_time(m[3]) for m in inner_marks) / ni
    tt = torch.tensor([tot], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    tot = float(tt.item())
    ms_per_step = tot / a.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- end to end through the public API: host buffers in, host arrays out -------------------------
    e2e = None
    if not a.no_e2e:
        import numpy as np
        x0 = torch.empty((n, D), dtype=torch.float32).pin_memory()
        x0.copy_(torch.randn(n, D) * 3.0)
        sc = mocat.scenarios.Rastrigin(dim=D, a=1.0, prior_std=3.0)

        def e2e_run(iters):
            smp = mocat.MetropolisedSMCSampler(mocat.Underdamped(stepsize=0.1), max_iter=iters, keep_history=False,
                                               check_every=16)
            t0 = time.perf_counter()
            out = mocat.run(sc, smp, n * world, random_key=a.seed, initial_state=mocat.cdict(value=x0.numpy()))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            d2h = sum(v.nbytes for v in out.__dict__.values() if isinstance(v, np.ndarray))
            return dt, len(out.temperature) - 1, d2h
        e2e_run(a.steps)                                             # warm-up: same configuration (engine pool, graphs)
        sync()
        dt, iters, d2h = e2e_run(a.steps)
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * n * iters / dt, "unit": "particle-steps/s",
               "h2d_bytes_per_step": n * D * 4 / max(iters, 1), "d2h_bytes_per_step": d2h / max(iters, 1),
               "iters": iters, "wall_s": dt, "api": "mocat_b200.run(scenario, MetropolisedSMCSampler, n, key, "
               "initial_state=cdict(value=<host ndarray>)) -> host cdict"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not a.no_cpu_baseline and world >= 1:
        n_s, k_s = 1_000_000, 5
        v, sec, workers = cpu_arm(n_s, k_s, 1, a.seed)
        cpu = {"value": v, "unit": "particle-steps/s", "cores": workers, "kind": "port",
               "sample": f"n={n_s} x {k_s} steps of the same workload; NumPy restatement of the reference "
                         f"(oracle/), {workers} processes; mocat's own JAX path cannot run here (jax absent)"}

    # roofline of the dominant kernel (by measured time) and of the move kernel.  Algorithmic bytes
    # (SURVEY 8d, C2): move 64 B/particle; tempering search 8 B x evaluations (w, l reads) + 12 B for the
    # weight update (read w, l; write w).
    evals = mean_search + 1.0
    bytes_temper = (8.0 * evals + 12.0) * n
    ach_move = ALGO_BYTES_MOVE * n / (t_move * 1e-3) / 1e9
    ach_temper = bytes_temper / (t_temper * 1e-3) / 1e9
    kernels = {
        "scan_cdf+ancestors (device-predicated)": {"ms": t_resample},
        "smc_move_kernel<Rastrigin,5,MALA>": {"ms": t_move, "algorithmic_bytes": ALGO_BYTES_MOVE * n,
                                              "achieved_gbs": ach_move, "frac": ach_move / hbm_peak},
        "temper_adapt_kernel<resident>": {"ms": t_temper, "algorithmic_bytes": bytes_temper,
                                          "achieved_gbs": ach_temper, "frac": ach_temper / hbm_peak,
                                          "mean_evaluations": evals},
    }
    dom = "temper_adapt_kernel<resident>" if t_temper >= t_move else "smc_move_kernel<Rastrigin,5,MALA>"
    line = {
        "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": a.steps,
        "warmup": nw, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_per_gpu": n, "dim": D, "parallelism": f"one population of {world * n} particles sharded over {world} GPUs "
                   "(peer-memory mailbox exchange + NVLink ancestor gather)" if world > 1 else "single GPU", "l2": "flushed between timed steps (512 MiB memset)",
                   "launch": "one CUDA-graph replay per step",
                   "state_bytes_resident": int(sum(t.numel() * t.element_size() for t in
                                                   (eng.xbuf[0], eng.xbuf[1], eng.lw, eng.lik, eng.up, eng.alpha,
                                                    eng.cdf, eng.anc)))},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"],
                     "peak": hbm_peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                     "traffic": NCU_DRAM_BYTES_PER_PARTICLE[dom] * n,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                     "ms_per_launch": kernels[dom]["ms"],
                     "note": "n=1e6: the whole state (68 MB) is smaller than L2 and every kernel is "
                             "latency/issue bound, not HBM bound; see DESIGN.md for the n=1e8 figures"},
        "kernels": kernels,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": (6 if world == 1 else 8) * a.steps,   # kernels per step: scan, strata hist, offsets scan, ancestors, move, temper (+ exchange, histogram sum when sharded)
        "clocks": clk,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Generate tests/golden/oracle_v3.npz.

The reference (mocat on JAX) cannot be imported in the build container (jax is absent: DESIGN.md section 2), so no
vectors could be produced by the reference itself.  These fixtures are seeded inputs together with the outputs of the
NumPy restatement (`oracle/`), which is pinned on the reference's own known-answer tests
(tests/test_oracle_kats.py).  They serve two purposes:
  * `-m "not gpu"`: the oracle must keep reproducing them (guards against drift of the checker);
  * `-m gpu`:       the CUDA path is compared against the stored numbers without importing `oracle/`.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import core, models, philox, svgd, smc, pf  # noqa: E402


def build():
    g = {}
    rng = np.random.default_rng(20241017)
    # Philox words / uniforms / normals
    gid = np.arange(16, dtype=np.uint64)
    raw = philox.raw(11, gid, 3, philox.P_MOVE, 2)
    g["philox_raw"] = np.stack([np.asarray(r, np.uint32) for r in raw])
    g["philox_normals"] = philox.normals(11, gid, 3, philox.P_MOVE, 5, dtype=np.float64)
    # LSE / ESS
    lw = (rng.standard_normal(1000) * 3).astype(np.float32)
    lw[::13] = -np.inf
    g["lse_lw"] = lw
    g["lse_out"] = np.asarray(core.lse_ess(lw), np.float64)
    # exact CDF and ancestors
    w = rng.random(5000).astype(np.float32) ** 3
    w[::11] = 0
    w = (w / w.astype(np.float64).sum()).astype(np.float32)
    cdf = core.cdf_from_weights(w, normalised=True)
    u = rng.random(5000)
    g["cdf_w"], g["cdf"], g["anc_u"] = w, cdf, u
    g["anc_multinomial"] = core.ancestors_multinomial(cdf, u).astype(np.int32)
    g["anc_systematic"] = core.ancestors_systematic(cdf, 0.37).astype(np.int32)
    # exact-rational systematic resampling on integer weights (csrc/resample_fused.cu, the engines' path)
    g["anc_exact_k0"] = np.array([0, 1, 0x9E3779B9, 0xFFFFFFFF], dtype=np.uint64)
    g["anc_systematic_exact"] = np.stack([core.ancestors_systematic_exact(core.integer_weights(w), int(k)).astype(np.int32)
                                          for k in g["anc_exact_k0"]])
    # pairwise normals of the Lorenz-96 kernel (csrc/pf_l96.cu)
    g["philox_normals_pairwise"] = philox.normals_pairwise(11, gid, 3, philox.P_MOVE, 8, dtype=np.float64)
    # quantile (jnp.quantile linear interpolation)
    v = rng.standard_normal(777).astype(np.float32)
    g["quant_v"] = v
    g["quant_q"] = np.array([0.0, 0.123, 0.5, 0.9, 1.0])
    g["quant_out"] = np.array([core.quantile_linear(v, q) for q in g["quant_q"]])
    # SVGD interaction + bandwidths
    X = rng.standard_normal((48, 3)) * 0.8 + 0.5
    G = rng.standard_normal((48, 3))
    X32, G32 = X.astype(np.float32), G.astype(np.float32)
    g["svgd_X"], g["svgd_G"] = X32, G32
    g["svgd_h"] = np.float32(1.3)
    g["svgd_phi"] = svgd.phi(X32.astype(np.float64), G32.astype(np.float64), np.float32(1.3))
    g["svgd_median_h"] = np.float64(svgd.median_bandwidth(X32))
    g["svgd_mean_h"] = np.float64(svgd.mean_bandwidth(X32))
    # logistic regression (config C4 target)
    A = rng.standard_normal((33, 7)).astype(np.float32)
    t = (rng.random(33) < 0.5).astype(np.float32)
    W = (rng.standard_normal((20, 7)) * 0.6).astype(np.float32)
    ul, gl = models.LogisticRegression(A, t).potential_and_grad(W)
    up, gp = models.IsoGaussianPrior(7, 0.0, 2.0).potential_and_grad(W)
    g["lr_A"], g["lr_t"], g["lr_W"] = A, t, W
    g["lr_U"], g["lr_G"] = up + 0.7 * ul, gp + 0.7 * gl
    # tempered SMC on Rastrigin d=2 (adaptive schedule, MALA, systematic): schedule and evidence of the oracle run
    sc = models.Rastrigin(2, 1.0)
    prior = models.IsoGaussianPrior(2, 0.0, 3.0)
    s = smc.TemperedSMC(prior, sc, n=512, seed=7, move="mala", stepsize=0.1, resampling="systematic",
                        normal_dtype=np.float32)
    chain = s.run()
    g["smc_beta"] = np.array([c["beta"] for c in chain])
    g["smc_log_z"] = np.array([c["log_norm_constant"] for c in chain])
    g["smc_ess"] = np.array([c["ess"] for c in chain])
    # bootstrap PF on the 1-D linear-Gaussian model of config C1 + Kalman log-likelihood
    lg = models.LinearGaussianSSM([0.0], [[1.0]], [[0.9]], [[0.5]], [[1.0]], [[0.3]])
    y = np.cumsum(rng.standard_normal(12))[:, None] * 0.3
    out = pf.BootstrapPF(lg, 2000, 5, resampling="systematic", normal_dtype=np.float32).run(y)
    g["pf_y"] = y
    g["pf_log_z"] = np.array([o["log_z"] for o in out])
    g["pf_ess"] = np.array([o["ess"] for o in out])
    g["kalman_loglik"] = np.float64(pf.kalman_filter(lg, y)[2])
    # Lorenz-96 d = 8 bootstrap filter: initial population and one step (fp64 RK4 flow, resampling every step)
    l96 = models.Lorenz96SSM(dim=8)
    _, yl = l96.simulate(2, np.random.default_rng(4), spinup=100)
    fl = pf.BootstrapPF(l96, 96, 13, ess_threshold=2.0, resampling="systematic")
    s0 = fl.init(yl[0])
    s1 = fl.step(s0, yl[1])
    g["l96_y"] = yl
    g["l96_x0"], g["l96_lw0"] = s0["x"], s0["lw"]
    g["l96_anc"], g["l96_x1"], g["l96_lw1"] = s1["ancestors"].astype(np.int32), s1["x"], s1["lw"]
    g["l96_log_z"] = np.array([s0["log_z"], s1["log_z"]])
    return g


if __name__ == "__main__":
    g = build()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v3.npz"), **g)
    print({k: (v.shape, str(v.dtype)) for k, v in g.items()})

"""Reductions, tempering search and resampling -- NumPy restatement.  TEST INFRASTRUCTURE.

Follows (all under /root/reference/mocat/src/):
  metrics.py:69-78          log_ess_log_weight / ess_log_weight
  utils.py:205-237          bisect  (regula falsi)
  transport/smc.py:303-326  MetropolisedSMCSampler.log_ess / next_temperature_adaptive
  transport/smc.py:61-71    SMCSampler.resample          (random.categorical + gather)
  ssm/filtering.py:196-199  _resample
  abc/smc.py:163-166        next_threshold_adaptive      (jnp.quantile, linear interpolation)
  abc/smc.py:94-98          adapt_stepsize_scaled_diag_cov (vmap(jnp.cov))
and jax.scipy.special.logsumexp (third party, unpinned; semantics restated from its
published source: m = max; m = 0 if not finite; log(sum(b*exp(a-m))) + m).
"""
import numpy as np

Q_BITS = 52                      # weights are quantised to multiples of 2^-52 before the fp64 cumsum
Q_SCALE = float(2 ** Q_BITS)
Q_INV = float(2.0 ** -Q_BITS)
Q_MARGIN = 1.0 + 2.0 ** -24      # un-normalised weights are scaled to sum to 1 + 6e-8 so the clamp at 1
                                 # (not the forced last element) closes the CDF; see DESIGN.md


# ----------------------------------------------------------------------------- LSE / ESS
def logsumexp(a, b=None):
    """jax.scipy.special.logsumexp semantics (call sites smc.py:160,214-215,288; metrics.py:74)."""
    a = np.asarray(a, dtype=np.float64)
    if a.size == 0:
        return -np.inf
    m = np.max(a)
    if not np.isfinite(m):
        m = 0.0
    with np.errstate(divide='ignore', invalid='ignore'):
        e = np.exp(a - m)
        s = np.sum(e if b is None else b * e)
        return float(np.log(s) + m)


def log_ess_log_weight(lw):
    """metrics.py:69-74: 2*LSE(w) - LSE(2w).  All -inf -> NaN (as in the reference)."""
    lw = np.asarray(lw, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        return 2.0 * logsumexp(lw) - logsumexp(2.0 * lw)


def ess_log_weight(lw):
    """metrics.py:77-78."""
    return float(np.exp(log_ess_log_weight(lw)))


def lse_ess(lw):
    """(lse, lse2, log_ess) -- what mb_lse_ess returns."""
    lw = np.asarray(lw, dtype=np.float64)
    l1 = logsumexp(lw)
    l2 = logsumexp(2.0 * lw)
    with np.errstate(invalid='ignore'):
        return l1, l2, 2.0 * l1 - l2


def lse_ess_tempered(lw, lik, dbeta):
    """log-ESS of  lw - dbeta*lik   (smc.py:303-309)."""
    return lse_ess(np.asarray(lw, np.float64) - float(dbeta) * np.asarray(lik, np.float64))


# ----------------------------------------------------------------------------- regula falsi
def bisect(fun, bounds, max_iter=1000, tol=1e-5):
    """utils.py:205-237 verbatim logic (the function is a secant/regula-falsi search).

    returns (bounds[2], evals[2], iters)."""
    b = [float(bounds[0]), float(bounds[1])]
    e = [float(fun(b[0])), float(fun(b[1]))]
    increasing = e[1] > e[0]                                   # :212
    it = 0
    while not (min(abs(e[0]), abs(e[1])) < tol or it >= max_iter
               or (e[0] < 0 and e[1] < 0) or (e[0] > 0 and e[1] > 0)):   # :214-218
        new_pos = b[0] - e[0] * (b[1] - b[0]) / (e[1] - e[0])  # :223
        new_eval = float(fun(new_pos))
        replace_upper = (new_eval > 0) if increasing else (new_eval < 0)   # :226
        if replace_upper:
            b, e = [b[0], new_pos], [e[0], new_eval]
        else:
            b, e = [new_pos, b[1]], [new_eval, e[1]]
        it += 1
    return b, e, it


def next_temperature_adaptive(lw, lik, beta, beta_max, ess0, retain=0.9, tol=1e-5, max_iter=1000):
    """transport/smc.py:311-326.  f(b') = log_ess(lw - (b'-beta)*lik) - log(ess0*retain);
    returns (beta_next, iters) with beta_next = bracket end with the smaller |f|."""
    lw = np.asarray(lw, np.float64)
    lik = np.asarray(lik, np.float64)
    log_target = np.log(ess0 * retain)
    f = lambda x: log_ess_log_weight(lw - (x - beta) * lik) - log_target
    b, e, it = bisect(f, [beta, beta_max], max_iter=max_iter, tol=tol)
    # jnp.argmin(jnp.abs(evals)): first index on ties, NaN counts as the minimum
    ae = np.abs(np.array(e))
    return b[int(np.argmin(ae))], it


# ----------------------------------------------------------------------------- resampling
def quantise_weights(w, scale=None):
    """Exact-fp64 convention (DESIGN.md "Resampling"): q_i = rint(w_i*scale) * 2^-52 with
    scale = 2^52 / sum (or 2^52 when the weights are already normalised: scale=None).

    All q_i are multiples of 2^-52 and sum to < 2, so EVERY fp64 partial sum is exact: the
    cumsum is associative, monotone and identical for any scan order or GPU sharding.  For
    fp32 weights >= 2^-28 the quantisation is the identity."""
    w = np.asarray(w).astype(np.float64)
    s = Q_SCALE if scale is None else float(scale)
    return np.rint(w * s) * Q_INV


def cdf_from_weights(w, normalised=True):
    """Inclusive fp64 cumsum of the quantised weights; clamped to 1 and last forced to 1.0.

    normalised=False: w are un-normalised (e.g. exp(lw - max)); scale = 2^52 / sum_fp64(w),
    where the sum is the exact fp64 sequential sum mirrored by the device reduction to
    within its stated tolerance -- tests feed the device's own `sum` back in via
    cdf_from_weights_scaled for bit-exactness."""
    w = np.asarray(w)
    if normalised:
        q = quantise_weights(w)
    else:
        q = quantise_weights(w, Q_SCALE * Q_MARGIN / float(np.sum(w.astype(np.float64))))
    return finish_cdf(np.cumsum(q))


def cdf_from_weights_scaled(w, scale):
    return finish_cdf(np.cumsum(quantise_weights(w, scale)))


def finish_cdf(c):
    c = np.minimum(c, 1.0)
    if c.size:
        c[-1] = 1.0
    return c


def cdf_from_log_weights(lw):
    """cdf of softmax(lw): e = exp(lw - max) evaluated in fp32 (as the device does), then
    cdf_from_weights(e, normalised=False).  Matches in law `random.categorical(key, lw)`
    (smc.py:65-67, filtering.py:199)."""
    lw = np.asarray(lw, dtype=np.float32)
    m = np.max(lw)
    if not np.isfinite(m):
        m = np.float32(0.0)
    with np.errstate(invalid='ignore'):
        e = np.exp((lw - m).astype(np.float32)).astype(np.float32)
    e = np.where(np.isnan(e), np.float32(0), e)
    return cdf_from_weights(e, normalised=False)


def systematic_uniforms(n_out, u0):
    """u_i = (i + u0) / n_out in fp64 (one add, one divide), i = 0..n_out-1."""
    return (np.arange(n_out, dtype=np.float64) + float(u0)) / float(n_out)


def ancestors_from_uniforms(cdf, u):
    """a_i = min{ j : cdf[j] > u_i } = searchsorted(cdf, u, side='right'), clipped to n-1."""
    cdf = np.asarray(cdf, dtype=np.float64)
    a = np.searchsorted(cdf, np.asarray(u, dtype=np.float64), side='right')
    return np.minimum(a, cdf.shape[0] - 1).astype(np.int64)


def ancestors_systematic(cdf, u0, n_out=None):
    n_out = len(cdf) if n_out is None else n_out
    return ancestors_from_uniforms(cdf, systematic_uniforms(n_out, u0))


def ancestors_multinomial(cdf, u):
    return ancestors_from_uniforms(cdf, u)


# ----------------------------------------------------------------------------- exact-rational systematic resampling
# Convention of csrc/resample_fused.cu (the particle filter's production resampler).  The reference only fixes the LAW
# of the ancestors (`random.categorical`, transport/smc.py:65-67, ssm/filtering.py:199); the device evaluates systematic
# resampling on integer weights in exact rational arithmetic so that no scan order / GPU sharding can change a bit:
#     e_i = rint(w_i * 2^K)  (uint64),  K = min(40, 63 - ceil(log2 n_total)),  C_j = sum_{i<=j} e_i,  S = C_{n-1},
#     u0 = k0 / 2^32,   a_i = min{ j : (i + u0) / n_out < C_j / S }.
def rs_scale_bits(n_total):
    lg = 0
    while (1 << lg) < int(n_total):
        lg += 1
    return min(40, 63 - lg)


def integer_weights(w, n_total=None):
    """e_i = rint(float32(w_i) * 2^K) as uint64 (linear weights <= 1; NaN / negative -> 0, as __float2ull_rn)."""
    w = np.asarray(w, dtype=np.float32)
    K = rs_scale_bits(w.shape[0] if n_total is None else n_total)
    v = w.astype(np.float64) * float(2 ** K)                 # exact: power-of-two scaling of an fp32 value
    v = np.where(np.isnan(v) | (v < 0), 0.0, v)
    return np.rint(v).astype(np.uint64)


def integer_weights_log(lw, n_total=None):
    """log mode: w_i = exp(lw_i - max lw) evaluated in fp32 like the device (the device uses the MUFU ex2 approximation,
    so individual e_i may differ in their last bits: parity of the ancestors from log-weights is statistical, the
    bit-exact contract is the linear mode)."""
    lw = np.asarray(lw, dtype=np.float32)
    m = np.max(lw)
    if not np.isfinite(m):
        m = np.float32(0.0)
    with np.errstate(invalid='ignore', over='ignore'):
        e = np.exp((lw - m).astype(np.float32)).astype(np.float32)
    return integer_weights(e, n_total)


def systematic_counts_exact(C, S, n_out, k0):
    """c_j = #{ i in [0, n_out) : (i * 2^32 + k0) * S < C_j * n_out * 2^32 } with Python integers (exact).
    Vectorised through an extended-precision estimate; every estimate within 1e-6 of an integer is settled exactly."""
    C = np.asarray(C, dtype=np.uint64)
    S, n_out, k0 = int(S), int(n_out), int(k0)
    if S == 0:
        return np.zeros(C.shape, dtype=np.int64)
    t = C.astype(np.longdouble) * (np.longdouble(n_out) / np.longdouble(S)) - np.longdouble(k0) / np.longdouble(2 ** 32)
    c = np.clip(np.ceil(t), 0, n_out).astype(np.int64)
    fr = t - np.floor(t)
    doubt = np.nonzero((fr < 1e-6) | (fr > 1 - 1e-6) | (C.astype(np.float64) >= float(S)))[0]
    for j in doubt:
        num = int(C[j]) * n_out * (1 << 32) - k0 * S          # i < num / (S 2^32)
        cj = 0 if num <= 0 else -((-num) // (S << 32))        # ceil division
        c[j] = min(max(cj, 0), n_out)
    return c


def ancestors_systematic_exact(e, k0, n_out=None):
    """ancestors of exact-rational systematic resampling from integer weights e (uint64) and the 32-bit offset k0.
    All-zero weights: every output takes the last particle (legacy convention cdf[n-1] = 1)."""
    e = np.asarray(e, dtype=np.uint64)
    n = e.shape[0]
    n_out = n if n_out is None else int(n_out)
    C = np.cumsum(e, dtype=np.uint64)
    S = int(C[-1])
    if S == 0:
        return np.full(n_out, n - 1, dtype=np.int64)
    c = systematic_counts_exact(C, S, n_out, k0)
    counts = np.diff(np.concatenate([[0], c]))
    return np.repeat(np.arange(n, dtype=np.int64), counts)


def strata_count(n_total_out):
    """B = largest power of two <= n/16 (at least 1, at most 2^24): ~16-32 outputs per stratum"""
    B = 1
    while B * 32 <= n_total_out and B < (1 << 24):
        B <<= 1
    return B


def ancestors_multinomial_stratified(cdf, seed, step, n_out=None):
    """Stratified-exact multinomial (mirrors csrc/resample.cu strata_hist/ancestors_sorted): the law of n iid
    uniforms = (histogram of first-stage uniforms over B equal strata) + (fresh second-stage uniforms inside each
    stratum); output g takes stratum s(g) = upper_bound(offsets, g) - 1 and u_g = (s + v_g)/B."""
    from . import philox
    n_out = len(cdf) if n_out is None else int(n_out)
    B = strata_count(n_out)
    g = np.arange(n_out, dtype=np.uint64)
    u1 = philox.uniform53(seed, g, step, philox.P_RESAMPLE, 0)
    hist = np.bincount((u1 * B).astype(np.int64), minlength=B)
    offsets = np.concatenate([[0], np.cumsum(hist)])
    s = np.searchsorted(offsets, np.arange(n_out), side='right') - 1
    v = philox.uniform53(seed, g, step, philox.P_RESAMPLE, 1)
    u = (s.astype(np.float64) + v) / float(B)
    return ancestors_from_uniforms(cdf, u), u


def categorical_gumbel(rng, lw, n_out):
    """Faithful restatement of jax.random.categorical(key, lw, shape=(n_out,)):
    argmax(Gumbel(n_out, n) + lw) -- O(n_out * n) draws (smc.py:65, filtering.py:199).
    Only feasible for n <= ~2e4; used as the 'faithful' CPU baseline."""
    lw = np.asarray(lw, dtype=np.float32)
    n = lw.shape[0]
    out = np.empty(n_out, dtype=np.int64)
    chunk = max(1, int(2 ** 24 // max(n, 1)))
    for s in range(0, n_out, chunk):
        e = min(n_out, s + chunk)
        g = rng.gumbel(size=(e - s, n)).astype(np.float32)
        out[s:e] = np.argmax(g + lw[None, :], axis=1)
    return out


def gather_state(cols, anc):
    """cdict.__getitem__ (core.py:46-56): every array field indexed along axis 0."""
    return [np.asarray(c)[anc] for c in cols]


# ----------------------------------------------------------------------------- ABC helpers
def quantile_linear(v, q):
    """jnp.quantile default (linear interpolation), as used at abc/smc.py:166."""
    v = np.sort(np.asarray(v, dtype=np.float64))
    n = v.shape[0]
    pos = float(q) * (n - 1)
    lo = int(np.floor(pos))
    hi = int(np.ceil(pos))
    lo = min(max(lo, 0), n - 1)
    hi = min(max(hi, 0), n - 1)
    frac = pos - np.floor(pos)
    return float(v[lo] * (1.0 - frac) + v[hi] * frac)


def colstats(x):
    """per-dimension mean and ddof=1 variance over ALL particles (abc/smc.py:97 vmap(jnp.cov))."""
    x = np.asarray(x, dtype=np.float64)
    return x.mean(axis=0), x.var(axis=0, ddof=1)


# ----------------------------------------------------------------------------- driver
def while_loop_stacked(cond_fun, body_fun, init_carry, max_iter=1000):
    """utils.py:159-202: run body while cond, stacking the state part of the carry."""
    carry = init_carry
    stack = []
    it = 0
    while it < max_iter and cond_fun(*carry):
        carry = body_fun(*carry)
        stack.append(carry[0])
        it += 1
    return stack, carry

"""CPU oracle for the mocat particle-population hot path.  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the algorithms in SamDuffield/mocat v0.2.6
(`/root/reference`, pure Python on JAX) that lie on the hot path named by
BASELINE.json: tempered SMC (`mocat/src/transport/smc.py`), the bootstrap particle
filter (`mocat/src/ssm/filtering.py`), SMC-ABC (`mocat/src/abc/smc.py`) and the SVGD
interaction (`mocat/src/transport/svgd.py`).  Every function cites the reference
file:line it follows.

Who may import it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` -- and there only as the *checker* (or the timed
CPU baseline), never as part of the shipped product path.  `mocat_b200/` must never
import `oracle`; a test enforces that.

Parity status.  The reference needs `jax`/`jaxlib` (`/root/reference/setup.py:15-16`),
which are not installed and not installable (no network, no wheel).  The oracle is pinned
(a) on OUTPUTS OF THE REFERENCE ITSELF: its own, unmodified source executed under a NumPy
stand-in for jax (`tests/golden/jaxshim/`, `tests/golden/make_reference_golden.py` and
`make_reference_runs.py`) -- deterministic functions to fp64 round-off
(`tests/test_reference_golden_cpu.py`), whole sampler runs through their statistics, SVGD
exactly (`tests/test_reference_runs_cpu.py`) -- and (b) on the reference's own
known-answer tests (leapfrog `mocat/src/tests/test_utils.py:137-163`, Gaussian
kernel `tests/test_kernels.py:16-29`, `gaussian_potential` vs scipy
`tests/test_utils.py:37-118`, `bisect` `:178-194`, `while_loop_stacked` `:166-175`,
tempered-SMC log-normalising-constant `tests/test_transport.py:49-55,121-163`, SVGD
moments `:65-118`, bootstrap-PF coverage `tests/test_ssm.py:33-44`) -- see
`tests/test_oracle_*.py`.  What is NOT pinned bit-wise, because the reference's own
tests do not pin it: the JAX threefry random streams, `random.categorical`
(statistical only), `odeint` for Lorenz-96 (never tested upstream), `jnp.quantile`
/`jnp.median`/adagrad (end-to-end moments only).  For those the oracle restates the
published algorithm of the third-party dependency (jax, unpinned in `setup.py`) and
DESIGN.md lists the conventions we had to fix ourselves (ancestor convention,
log-evidence for the particle filter, fixed-step RK4 flow, Philox streams).
"""

from . import philox, core, models, mcmc, smc, pf, abc, svgd  # noqa: F401

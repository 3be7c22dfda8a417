"""Fused exact-integer systematic resampler (csrc/resample_fused.cu) against oracle/core.py: bit-exact ancestors
given weights and u0, edge cases (empty tiles, zero weights, collapsed weights -> heavy-tile pass, ragged sizes),
P-independence through single-GPU VIRTUAL RANKS (the sharded code path with every `peer` pointer on one device),
and the headline size n = 1e8."""
import ctypes as C

import numpy as np
import pytest

from oracle import core

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(lib):
    import torch
    from mocat_b200 import _lib, engine
    return torch, _lib, engine, lib


def _ws(torch, lib, n):
    return torch.zeros((int(lib.dll.mb_rs_workspace_bytes(n)) + 7) // 8, dtype=torch.int64, device="cuda")


def fused_ancestors(E, w, k0, n_out=None):
    """single shard, linear mode"""
    torch, l, e, lib = E
    n = len(w)
    wd = torch.as_tensor(np.asarray(w, np.float32), device="cuda")
    ws = _ws(torch, lib, n)
    anc = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    lib.call("mb_rs_tile_sums", lib.ctx(), l.ptr(ws), l.ptr(wd), n, n, 0, None, 1, l.stream())
    lib.call("mb_rs_ancestors", lib.ctx(), l.ptr(ws), l.ptr(wd), n, n, 0, None, 1, int(k0), None, None, l.ptr(anc),
             l.stream())
    return anc.cpu().numpy().astype(np.int64)


def virtual_rank_ancestors(E, w, k0, world):
    """the SHARDED path on one GPU: `world` shards of equal size.  Pass A per rank, the totals, pass B per rank (own source
    tiles; ancestors written through the anc_peers table = slices of one device array; heavy tiles recorded), then --
    after what is a barrier on real ranks -- pass C per rank (every rank fills its own share of ALL ranks' heavy tiles,
    reading their weights / records through lw_peers / ws_peers)"""
    torch, l, e, lib = E
    n = len(w)
    assert n % world == 0
    nl = n // world
    wd = torch.as_tensor(np.asarray(w, np.float32), device="cuda")
    anc = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    wss = [_ws(torch, lib, nl) for _ in range(world)]
    for r in range(world):
        lib.call("mb_rs_tile_sums", lib.ctx(), l.ptr(wss[r]), l.ptr(wd[r * nl:]), nl, n, 0, None, 1, l.stream())
    totals = torch.stack([ws[0] for ws in wss]).contiguous()             # uint64 bit patterns in int64
    shards = []
    for r in range(world):
        sh = l.Shard()
        sh.rank, sh.world, sh.n_local, sh.n_total = r, world, nl, n
        for q in range(world):
            sh.anc_peers[q] = anc[q * nl:].data_ptr()
            sh.lw_peers[q] = wd[q * nl:].data_ptr()
            sh.ws_peers[q] = wss[q].data_ptr()
        shards.append(sh)
    for r in range(world):
        lib.call("mb_rs_ancestors", lib.ctx(), l.ptr(wss[r]), l.ptr(wd[r * nl:]), nl, n, 0, None, 1, int(k0),
                 l.ptr(totals), C.byref(shards[r]), l.ptr(anc[r * nl:]), l.stream())
    for r in range(world):
        lib.call("mb_rs_heavy", lib.ctx(), l.ptr(wss[r]), l.ptr(wd[r * nl:]), nl, n, 0, None, 1, int(k0),
                 l.ptr(totals), C.byref(shards[r]), l.ptr(anc[r * nl:]), l.stream())
    return anc.cpu().numpy().astype(np.int64)


def _weights(kind, n, rng):
    if kind == "uniform":
        w = rng.random(n)
    elif kind == "lognormal":                       # wide dynamic range, many weights below the integer resolution
        w = np.exp(rng.standard_normal(n) * 6.0)
    elif kind == "sparse":                          # mostly zero: empty tiles, long runs without offspring
        w = rng.random(n) * (rng.random(n) < 0.01)
        w[rng.integers(n)] = 1.0
    elif kind == "onehot":                          # one particle owns everything: heavy-tile pass
        w = np.zeros(n)
        w[int(0.37 * n)] = 1.0
    elif kind == "spikes":                          # a few particles own almost everything
        w = rng.random(n) * 1e-6
        w[rng.integers(n, size=5)] = rng.random(5) + 0.5
    elif kind == "equal":
        w = np.ones(n)
    w = np.asarray(w, np.float64)
    return (w / w.max()).astype(np.float32)         # linear mode contract: weights <= 1


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 10_000, 1_000_003])
@pytest.mark.parametrize("kind", ["uniform", "lognormal", "sparse", "onehot", "spikes", "equal"])
def test_fused_systematic_bit_exact(E, n, kind):
    rng = np.random.default_rng(n * 7 + len(kind))
    w = _weights(kind, n, rng)
    e = core.integer_weights(w)
    for k0 in (0, 1, 0x80000000, 0xffffffff, int(rng.integers(1 << 32))):
        got = fused_ancestors(E, w, k0)
        ref = core.ancestors_systematic_exact(e, k0)
        assert np.array_equal(got, ref), (n, kind, k0, int(np.sum(got != ref)))


def test_fused_all_zero_weights(E):
    n = 5000
    got = fused_ancestors(E, np.zeros(n, np.float32), 123)
    assert np.all(got == n - 1)                      # legacy convention: cdf[n-1] = 1


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind", ["uniform", "lognormal", "onehot", "spikes"])
def test_virtual_ranks_match_single_shard(E, world, kind):
    """SURVEY 8e 'independent of P': the sharded resampler (tile sums per rank, totals exchange, outputs written to
    the owning rank) gives the SAME BITS as one shard and as the oracle."""
    n = 8 * 32 * 1021                                # 32672 particles per shard at world = 8 (>= 8192)
    rng = np.random.default_rng(world * 100 + len(kind))
    w = _weights(kind, n, rng)
    k0 = int(rng.integers(1 << 32))
    ref = core.ancestors_systematic_exact(core.integer_weights(w), k0)
    one = fused_ancestors(E, w, k0)
    many = virtual_rank_ancestors(E, w, k0, world)
    assert np.array_equal(one, ref)
    assert np.array_equal(many, ref), int(np.sum(many != ref))


def test_fused_log_mode_matches_oracle_statistically(E):
    """log mode evaluates exp(lw - max) with the MUFU approximation: individual integer weights can differ in their last
    bits from the fp32 NumPy exp of the oracle, so the ancestors agree up to a tiny fraction of boundary flips"""
    torch, l, e, lib = E
    n = 400_000
    rng = np.random.default_rng(5)
    lw = (rng.standard_normal(n) * 3.0 - 50.0).astype(np.float32)
    lwd = torch.as_tensor(lw, device="cuda")
    ctl = e.ControlBlock()
    rec = np.zeros(1, dtype=l.CONTROL_DTYPE)[0]
    rec['wmax'], rec['resample'], rec['seed'], rec['iter'] = float(lw.max()), 1, 11, 3
    ctl.write(rec)
    ws = _ws(torch, lib, n)
    anc = torch.empty(n, dtype=torch.int32, device="cuda")
    lib.call("mb_rs_tile_sums", lib.ctx(), l.ptr(ws), l.ptr(lwd), n, n, 1, l.ptr(ctl.t), 0, l.stream())
    lib.call("mb_rs_ancestors", lib.ctx(), l.ptr(ws), l.ptr(lwd), n, n, 1, l.ptr(ctl.t), 0, -1, None, None, l.ptr(anc),
             l.stream())
    from oracle import philox
    k0 = int(philox.uniform32(11, np.zeros(1, np.uint64), 4, philox.P_RESAMPLE)[0])
    ref = core.ancestors_systematic_exact(core.integer_weights_log(lw), k0)
    got = anc.cpu().numpy()
    assert np.mean(got != ref) < 2e-3
    assert np.max(np.abs(got - ref)) <= 4             # a flipped boundary moves an output to a neighbouring particle with offspring


def test_fused_predicated_on_control_block(E):
    torch, l, e, lib = E
    n = 10_000
    lwd = torch.zeros(n, device="cuda")
    ctl = e.ControlBlock()
    rec = np.zeros(1, dtype=l.CONTROL_DTYPE)[0]
    rec['resample'] = 0
    ctl.write(rec)
    ws = _ws(torch, lib, n)
    anc = torch.full((n,), -7, dtype=torch.int32, device="cuda")
    lib.call("mb_rs_tile_sums", lib.ctx(), l.ptr(ws), l.ptr(lwd), n, n, 1, l.ptr(ctl.t), 0, l.stream())
    lib.call("mb_rs_ancestors", lib.ctx(), l.ptr(ws), l.ptr(lwd), n, n, 1, l.ptr(ctl.t), 0, 5, None, None, l.ptr(anc),
             l.stream())
    assert np.all(anc.cpu().numpy() == -7)           # ctl->resample == 0: nothing ran


# ------------------------------------------------------------------------------------------------ headline size
def test_fused_systematic_1e8(E):
    """n = 1e8 (config C3's population): bit-exact against the oracle evaluated in chunks, and the number of ancestors
    that a plain fp64 `np.cumsum` + `searchsorted` of the same weights would place differently (quantify
    the departure from the literal fp64 cumsum)."""
    n = 100_000_000
    rng = np.random.default_rng(2026)
    w = rng.random(n, dtype=np.float32)
    w[rng.integers(n, size=1000)] = 0.0
    k0 = 0x9e3779b9
    got = fused_ancestors(E, w, k0)
    e = core.integer_weights(w)
    C = np.cumsum(e, dtype=np.uint64)
    S = int(C[-1])
    # ancestors are sorted and a_i = j  <=>  c_{j-1} <= i < c_j: verify through the counts at the ancestors' boundaries
    assert np.all(np.diff(got) >= 0) and got[0] >= 0 and got[-1] <= n - 1
    counts = np.bincount(got, minlength=n)
    c_dev = np.cumsum(counts)
    step = 10_000_000
    for lo in range(0, n, step):
        ref = core.systematic_counts_exact(C[lo:lo + step], S, n, k0)
        assert np.array_equal(c_dev[lo:lo + step], ref), lo
    # departure from the plain floating-point evaluation
    cdf = C.astype(np.float64) / float(S)
    flips = 0
    for lo in range(0, n, step):
        u = (np.arange(lo, min(n, lo + step), dtype=np.float64) + k0 / 2.0 ** 32) / n
        a = np.minimum(np.searchsorted(cdf, u, side='right'), n - 1)
        flips += int(np.sum(a != got[lo:lo + step]))
    print(f"n=1e8: {flips} ancestors differ between exact-rational and fp64 searchsorted evaluation")
    assert flips <= 50                                # ~ n^2 * 2^-53 boundary cases, each by one index

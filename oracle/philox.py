"""Philox4x32-10 counter RNG and the uniform/normal conversions shared with the CUDA kernels.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference draws its randomness from
`jax.random` (threefry2x32; e.g. `mocat/src/transport/smc.py:82`,
`mocat/src/ssm/filtering.py:294`, `mocat/src/mcmc/standard_mcmc.py:54-56,120-122`,
`mocat/src/mcmc/metropolis.py:60-61`).  jax is absent and the reference's tests pin
those streams only statistically ("parity unpinned" at the RNG boundary), so the B200
build defines its own counter-based stream (north star: "Philox RNG") and this module
mirrors it bit-for-bit so that CPU oracle and GPU kernels see the SAME random numbers:

    counter = (gid_lo, gid_hi, step, slot)      key = (seed_lo, seed_hi)
    slot    = (purpose << 20) | index

`gid` is the GLOBAL particle (or output-slot) index, so results do not depend on how
particles are sharded over GPUs.  Philox4x32-10 is Salmon et al. (SC'11); constants as
in Random123 / cuRAND.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)

# purposes (must match mocat_b200/csrc/rng.cuh)
P_INIT = 0      # initial / prior sample normals
P_MOVE = 1      # proposal normals + accept uniform of the move kernels
P_RESAMPLE = 2  # resampling uniforms (gid = output slot, or 0 for systematic u0)
P_SIM = 3       # simulator draws (ABC likelihood_sample)
P_OBS = 4       # data simulation (host-side helper streams)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for r in range(10):
            if r > 0:
                k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
                k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
    return c0, c1, c2, c3


def slot_of(purpose, index):
    return np.uint32((int(purpose) << 20) | int(index))


def raw(seed, gid, step, purpose, index):
    """4 raw uint32 words for every gid (array of uint64/int)."""
    gid = np.asarray(gid, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return philox4x32_10((gid & MASK).astype(np.uint32), (gid >> np.uint64(32)).astype(np.uint32),
                         np.uint32(int(step) & 0xFFFFFFFF), slot_of(purpose, index),
                         seed & 0xFFFFFFFF, seed >> 32)


def u24(x):
    """[0,1) uniform with 24 bits: (x >> 8) * 2^-24, exact in fp32."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def u_open(x):
    """(0,1] uniform: (float32(x) + 0.5f) * 2^-32 with fp32 round-to-nearest at each step."""
    xf = np.asarray(x, dtype=np.uint32).astype(np.float32)
    return (xf + np.float32(0.5)) * np.float32(2.0 ** -32)


def u53(xa, xb):
    """[0,1) fp64 uniform from two words: ((xa >> 5) * 2^26 + (xb >> 6)) * 2^-53."""
    a = (np.asarray(xa, dtype=np.uint32) >> np.uint32(5)).astype(np.float64)
    b = (np.asarray(xb, dtype=np.uint32) >> np.uint32(6)).astype(np.float64)
    return (a * 67108864.0 + b) * (2.0 ** -53)


def box_muller(xa, xb, dtype=np.float64):
    """(z0, z1) = sqrt(-2 ln u_open(xa)) * (cos, sin)(2 pi u24(xb)).

    The uniforms are the fp32 values the device sees; the transcendental part is evaluated
    in `dtype` (fp64 by default: the device uses fp32 fast intrinsics and is compared
    within tolerance)."""
    u1 = u_open(xa).astype(dtype)
    u2 = u24(xb).astype(dtype)
    r = np.sqrt(-2.0 * np.log(u1))
    th = 2.0 * np.pi * u2
    return r * np.cos(th), r * np.sin(th)


def normals(seed, gid, step, purpose, count, index0=0, dtype=np.float64):
    """`count` standard normals per gid -> array (len(gid), count).

    Normal j comes from slot index0 + j//4, words (0,1) for j%4 in {0,1} and (2,3) for {2,3}."""
    gid = np.atleast_1d(np.asarray(gid, dtype=np.uint64))
    out = np.empty((gid.shape[0], 4 * ((count + 3) // 4)), dtype=dtype)
    for s in range((count + 3) // 4):
        x0, x1, x2, x3 = raw(seed, gid, step, purpose, index0 + s)
        out[:, 4 * s + 0], out[:, 4 * s + 1] = box_muller(x0, x1, dtype)
        out[:, 4 * s + 2], out[:, 4 * s + 3] = box_muller(x2, x3, dtype)
    return out[:, :count]


def uniforms24(seed, gid, step, purpose, count, index0=0):
    """`count` fp32 [0,1) uniforms per gid (4 per slot) -> (len(gid), count) float32."""
    gid = np.atleast_1d(np.asarray(gid, dtype=np.uint64))
    out = np.empty((gid.shape[0], 4 * ((count + 3) // 4)), dtype=np.float32)
    for s in range((count + 3) // 4):
        xs = raw(seed, gid, step, purpose, index0 + s)
        for k in range(4):
            out[:, 4 * s + k] = u24(xs[k])
    return out[:, :count]


def uniform53(seed, gid, step, purpose, index=0):
    """One fp64 [0,1) uniform per gid from words (0,1) of the slot."""
    x0, x1, _, _ = raw(seed, gid, step, purpose, index)
    return u53(x0, x1)


def normals_pairwise(seed, gid, step, purpose, count, dtype=np.float64):
    """`count` standard normals per gid with the PAIRWISE convention of csrc/pf_l96.cu: particles 2m and 2m+1 share
    the Philox stream of their pair id m; coordinate k comes from slot k//2, words (0,1) for even k and (2,3) for odd
    k; the cos branch of Box-Muller goes to the even particle, the sin branch to the odd one."""
    gid = np.atleast_1d(np.asarray(gid, dtype=np.uint64))
    pair = gid >> np.uint64(1)
    odd = (gid & np.uint64(1)).astype(bool)
    out = np.empty((gid.shape[0], count), dtype=dtype)
    for c in range((count + 1) // 2):
        x0, x1, x2, x3 = raw(seed, pair, step, purpose, c)
        zc, zs = box_muller(x0, x1, dtype)
        out[:, 2 * c] = np.where(odd, zs, zc)
        if 2 * c + 1 < count:
            zc, zs = box_muller(x2, x3, dtype)
            out[:, 2 * c + 1] = np.where(odd, zs, zc)
    return out


def uniform32(seed, gid, step, purpose, index=0):
    """word 0 of the slot as an integer in [0, 2^32): the systematic offset u0 = k0 / 2^32 of resample_fused.cu"""
    x0, _, _, _ = raw(seed, gid, step, purpose, index)
    return x0.astype(np.uint64)

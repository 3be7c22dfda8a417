"""CPU tests (gloo, world_size 2) of the host-side sharding logic in mocat_b200/parallel.py: shard ranges,
owner arithmetic, CDF offsets and the handle exchange used to connect the peer-mapped mailboxes.  The kernels
themselves need GPUs (tests/test_gpu_multi.py)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKER = r'''
import os, sys, numpy as np, torch.distributed as dist
sys.path.insert(0, os.environ["MB_ROOT"])
from mocat_b200 import parallel
from oracle import core
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# handle exchange (what ShardContext does with the 64-byte CUDA IPC handles)
mine = bytes([rank]) * 64
got = parallel.all_gather_bytes(mine)
assert got == [bytes([r]) * 64 for r in range(world)], got
# sharded exact CDF == single-rank CDF: every rank scans its shard relative to 0, offsets = exclusive prefix of totals
n = 10_000
rng = np.random.default_rng(0)
w = rng.random(n).astype(np.float32) ** 3
w = (w / w.astype(np.float64).sum()).astype(np.float32)
g0, nl = parallel.shard_range(n, rank, world)
q = core.quantise_weights(w[g0:g0 + nl])
local = np.cumsum(q)
totals = [None] * world
dist.all_gather_object(totals, float(local[-1]))
off = parallel.global_cdf_offsets(totals)
glob = core.finish_cdf(off[rank] + local) if rank == world - 1 else np.minimum(off[rank] + local, 1.0)
ref = core.cdf_from_weights(w)[g0:g0 + nl]
assert np.array_equal(glob, ref), "sharded CDF differs from the single-rank CDF"
# owner arithmetic of global ancestors
anc = core.ancestors_systematic(core.cdf_from_weights(w), 0.3)
owner, idx = parallel.owner_of(anc, nl)
assert np.array_equal(owner * nl + idx, anc) and owner.max() < world
dist.barrier()
print("PARALLEL_CPU_OK", rank)
'''


def test_shard_helpers():
    from mocat_b200 import parallel
    assert parallel.shard_range(100, 3, 4) == (75, 25)
    owner, idx = parallel.owner_of(np.array([0, 24, 25, 99]), 25)
    assert owner.tolist() == [0, 0, 1, 3] and idx.tolist() == [0, 24, 0, 24]
    np.testing.assert_array_equal(parallel.global_cdf_offsets([0.25, 0.5, 0.25]), [0.0, 0.25, 0.75, 1.0])


def test_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MB_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29655", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("PARALLEL_CPU_OK") == 2

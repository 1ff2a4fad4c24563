"""N>1 host logic on CPU: world_size-2 gloo run of the theta-sharding / scalar all-reduce."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from spdepy_b200 import parallel as par
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    calls = []
    def fun(t):
        calls.append(1)
        return float((t ** 2).sum()), 2 * t
    thetas = np.arange(15, dtype=float).reshape(5, 3)
    likes, jacs = par.evaluate_thetas(fun, thetas)
    assert len(calls) == len(par.shard_indices(5)) and len(calls) in (2, 3)
    assert np.allclose(likes, (thetas ** 2).sum(axis=1)) and np.allclose(jacs, 2 * thetas)
    g = par.finite_difference_gradient(lambda t: float((t ** 3).sum()), np.array([1.0, 2.0, -1.0]), h=1e-4)
    assert np.allclose(g, 3 * np.array([1.0, 4.0, 1.0]), atol=1e-6)
    blocks = [par.column_block(10, r, world) for r in range(world)]
    assert sum(b.stop - b.start for b in blocks) == 10
    class Stub:                      # Model.sample contract: same seeded draw, only the requested columns
        def sample(self, n, simple, seed, cols):
            z = np.random.default_rng(seed).normal(size=6 * n).reshape(6, n)
            return 2.0 * z[:, cols] + 1.0
    full = par.sample_sharded(Stub(), 7, seed=3)
    ref = 2.0 * np.random.default_rng(3).normal(size=42).reshape(6, 7) + 1.0
    assert full.shape == (6, 7) and np.array_equal(full, ref)
    cols, mine = par.sample_sharded(Stub(), 7, seed=3, gather=False)
    assert np.array_equal(mine, ref[:, cols]) and mine.shape[1] in (3, 4)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""") % ROOT


def test_theta_sharding_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_single_process_defaults():
    import numpy as np
    from spdepy_b200 import parallel as par
    likes, jacs = par.evaluate_thetas(lambda t: (t.sum(), t), np.ones((3, 2)))
    assert np.allclose(likes, 2.0) and jacs.shape == (3, 2)
    assert par.shard_indices(5, 1, 2) == [1, 3]

"""Multi-GPU sharding of the hot path: one process per GPU, replicated factorisations.

The reference's only parallelism is launching independent Python processes in tmux, one theta (or one
stochastic-gradient repeat) per process (``examples/server/run_grad_test.sh:7-40``,
``examples/server/grad_test.py:10-22``).  The natural shards are therefore independent theta evaluations
(finite-difference checks, line searches, sweeps) and column blocks of samples / right-hand sides.  Each
rank owns its own factorisation; the only data-path collective is ONE all-reduce of the ``npar+1``
likelihood / gradient scalars per batch (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items: int, rank: int | None = None, world: int | None = None) -> list[int]:
    """Round-robin ownership of ``n_items`` independent units."""
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    return list(range(rank, n_items, world))


def _buffer_device():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def evaluate_thetas(fun, thetas):
    """Evaluate ``fun(theta) -> (like, jac)`` for every row of ``thetas``, sharded round-robin over the
    ranks; returns ``(likes, jacs)`` complete on every rank after a single SUM all-reduce."""
    thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    m, npar = thetas.shape
    buf = torch.zeros(m, npar + 1, dtype=torch.float64, device=_buffer_device())
    for i in shard_indices(m):
        like, jac = fun(thetas[i])
        buf[i, 0] = float(like)
        buf[i, 1:] = torch.as_tensor(np.asarray(jac, dtype=np.float64), device=buf.device)
    if _world()[1] > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    out = buf.cpu().numpy()
    return out[:, 0].copy(), out[:, 1:].copy()


def finite_difference_gradient(fun_value, theta, h=1e-3):
    """Central differences of ``fun_value(theta)`` with the 2*npar evaluations sharded over the ranks
    (the check of ``examples/server/grad_test.py:7-17``)."""
    theta = np.asarray(theta, dtype=np.float64)
    npar = theta.size
    pts = np.repeat(theta[None, :], 2 * npar, axis=0)
    for i in range(npar):
        pts[2 * i, i] += h
        pts[2 * i + 1, i] -= h
    vals, _ = evaluate_thetas(lambda t: (fun_value(t), np.zeros(npar)), pts)
    return (vals[0::2] - vals[1::2]) / (2 * h)


def column_block(n_cols: int, rank: int | None = None, world: int | None = None) -> slice:
    """Contiguous block of sample / right-hand-side columns owned by a rank."""
    r, w = _world()
    rank = r if rank is None else rank
    world = w if world is None else world
    per = (n_cols + world - 1) // world
    return slice(min(rank * per, n_cols), min((rank + 1) * per, n_cols))


def sample_sharded(model, n: int, seed: int, simple: bool = True, gather: bool = True):
    """``Model.sample(n, seed=seed)`` with the ``n`` columns split into contiguous blocks over the ranks: every rank
    factorises the same precision (replicated, as everywhere on this path), draws the same seeded ``z`` and solves only
    its own columns (``Model.sample(..., cols=...)``).  With ``gather`` one all-gather assembles the full block on
    every rank, otherwise each rank keeps ``(columns, block)``."""
    rank, world = _world()
    cols = column_block(n, rank, world)
    mine = np.asarray(model.sample(n=n, simple=simple, seed=seed, cols=cols), dtype=np.float64)
    if not gather:
        return cols, mine
    if world == 1:
        return mine
    per = (n + world - 1) // world
    pad = torch.zeros(mine.shape[0], per, dtype=torch.float64, device=_buffer_device())
    pad[:, :mine.shape[1]] = torch.as_tensor(mine, device=pad.device)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = np.empty((mine.shape[0], n))
    for r in range(world):
        c = column_block(n, r, world)
        out[:, c] = parts[r][:, :c.stop - c.start].cpu().numpy()
    return out

"""Host-side logic of the multi-GPU candidate sweep: how a candidate range is split across ranks and
how the per-rank winners are combined. The data path has no collective; the only exchange is one (value, index)
pair per rank (16 bytes), gathered with torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The reference's counterpart is the multi-start arg-max of FindGlobalSolution
(src/acquisition-function.cpp:122-153): every thread evaluates its own starts, then the best one is taken.
"""
from __future__ import annotations

import numpy as np

_M1, _M2, _M3, _M4 = (np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1B54A32D192ED03), np.uint64(0xBF58476D1CE4E5B9),
                      np.uint64(0x94D049BB133111EB))


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of [0, total): returns (first, count) of `rank`; the first total % world ranks get
    one extra candidate. Contiguity keeps the lowest-index tie-break meaningful across ranks."""
    if world <= 0 or not 0 <= rank < world or total < 0:
        raise ValueError("shard_range: bad arguments")
    base, rem = divmod(total, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def candidate_coords(seed: int, first: int, count: int, D: int) -> np.ndarray:
    """The counter-based candidate generator of csrc/common.cuh (candidate_coord) in numpy: D x count, column-major.
    Coordinate d of candidate i depends on (seed, i, d) only, so any split of a range yields the same points."""
    with np.errstate(over="ignore"):
        i = (np.arange(first, first + count, dtype=np.uint64) + np.uint64(1))[None, :]
        d = (np.arange(D, dtype=np.uint64) + np.uint64(1))[:, None]
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + _M1 * i + _M2 * d
        z = (z ^ (z >> np.uint64(30))) * _M3
        z = (z ^ (z >> np.uint64(27))) * _M4
        z = z ^ (z >> np.uint64(31))
    return np.asfortranarray((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))


def select_winner(pairs) -> tuple[float, int]:
    """pairs: (world, 2) array of per-rank (best value, global index); index < 0 marks a rank without a valid
    candidate (empty shard or all-NaN). Highest value wins, lowest index breaks ties, NaN never wins."""
    pairs = np.asarray(pairs, dtype=np.float64).reshape(-1, 2)
    ok = (pairs[:, 1] >= 0) & ~np.isnan(pairs[:, 0])
    if not ok.any():
        raise ValueError("select_winner: no rank has a valid candidate")
    p = pairs[ok]
    w = np.lexsort((p[:, 1], -p[:, 0]))[0]
    return float(p[w, 0]), int(p[w, 1])


def all_gather_winner(value: float, index: int, device=None, group=None) -> tuple[float, int]:
    """Every rank contributes its local winner; every rank returns the same global winner."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return select_winner([[value, index]])
    mine = torch.tensor([value, float(index)], dtype=torch.float64, device=device)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, mine, group=group)
    return select_winner(torch.stack(out).cpu().numpy())


def select_best_point(values, points) -> tuple[float, np.ndarray, int]:
    """values: (world,), points: (world, D): the refined maximiser of every rank (slsgp_acq_maximize on its own candidate
    range). Highest value wins, the lowest rank breaks ties, NaN never wins. Returns (value, point, winning rank)."""
    values = np.asarray(values, dtype=np.float64).reshape(-1)
    points = np.asarray(points, dtype=np.float64).reshape(len(values), -1)
    ok = ~np.isnan(values)
    if not ok.any():
        raise ValueError("select_best_point: every rank reported NaN")
    r = int(np.flatnonzero(ok)[np.argmax(values[ok])])  # argmax returns the first (lowest-rank) maximum
    return float(values[r]), points[r].copy(), r


def all_gather_best_point(value: float, point, device=None, group=None) -> tuple[float, np.ndarray, int]:
    """Multi-GPU form of slsgp_acq_maximize: every rank contributes (value, x) of its local maximiser, 8 (1 + D) bytes;
    every rank returns the same global one."""
    import torch
    import torch.distributed as dist

    point = np.asarray(point, dtype=np.float64).reshape(-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return select_best_point([value], [point])
    mine = torch.tensor(np.concatenate([[value], point]), dtype=torch.float64, device=device)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, mine, group=group)
    allv = torch.stack(out).cpu().numpy()
    return select_best_point(allv[:, 0], allv[:, 1:])

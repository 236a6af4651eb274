"""One-process-per-GPU sharding of the edge construction.

The path shards without any data-path exchange: the sorted distinct-barcode array is replicated, the rows
are dealt to ranks in tiles of BDG_ROW_TILE (boustrophedon, include/badger_b200.h), every rank emits the edges
whose smaller barcode it owns, and the per-rank lists are concatenated.  ``torch.distributed`` is only the
plumbing for that final concatenation (NCCL when the ranks hold GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import ops
from ._lib import ROW_TILE


def owner_of_tile(tile: int, nparts: int) -> int:
    m = tile % (2 * nparts)
    return m if m < nparts else 2 * nparts - 1 - m


def part_rows(n: int, part: int, nparts: int) -> np.ndarray:
    """Indices (into the sorted array) of the rows owned by `part`."""
    tiles = np.arange((n + ROW_TILE - 1) // ROW_TILE)
    m = tiles % (2 * nparts)
    own = np.where(m < nparts, m, 2 * nparts - 1 - m) == part
    rows = (tiles[own][:, None] * ROW_TILE + np.arange(ROW_TILE)[None, :]).reshape(-1)
    return rows[rows < n]


def part_pairs(n: int, part: int, nparts: int) -> int:
    """Number of unordered pairs decided by `part`: sum over its rows i of (n-1-i)."""
    rows = part_rows(n, part, nparts).astype(np.int64)
    return int(((n - 1) - rows).sum())


def edges_build_distributed(sorted_unique: np.ndarray, t: int, build_part=None, gather: bool = True):
    """Each rank of the default process group builds its part; with gather=True every rank returns the
    union (a, b, d), else only its own part.  build_part(sorted, t, part, nparts) defaults to the GPU
    operator ops.edges_build_part."""
    import torch
    import torch.distributed as dist
    if build_part is None:
        build_part = ops.edges_build_part
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    a, b, d = build_part(sorted_unique, t, rank, world)
    if world == 1 or not gather:
        return a, b, d
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_local = torch.tensor([a.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    packed = np.zeros((cap, 3), dtype=np.int64)                 # (a, b, d) rows, padded to the longest part
    packed[:a.size, 0], packed[:a.size, 1], packed[:a.size, 2] = a, b, d
    mine = torch.from_numpy(packed).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    rows = np.concatenate([p.cpu().numpy()[:n] for p, n in zip(parts, sizes)], axis=0)
    return rows[:, 0].astype(np.uint32), rows[:, 1].astype(np.uint32), rows[:, 2].astype(np.uint8)

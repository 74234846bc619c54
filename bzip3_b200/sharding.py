"""Host-side multi-GPU logic of the block codec: static block sharding and the one exchange the path has,
the ordered gather of variable-length compressed blocks to the writer rank.
Backend-agnostic (`dist` is torch.distributed, NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def blocks_of_rank(n_blocks: int, rank: int, world: int) -> list[int]:
    """Static work queue: block i belongs to rank i mod world (blocks are equal-sized except the last)."""
    return list(range(rank, n_blocks, world))


def gather_compressed(enc: dict, n_blocks: int, rank: int, world: int, dist, device="cuda") -> dict | None:
    """enc: {block index: bytes} of this rank.  Returns {block index: bytes} for all blocks on rank 0, None elsewhere.
    Sizes travel by all_gather (4 B per block); payloads are padded to the largest block and gathered."""
    import torch
    mine = blocks_of_rank(n_blocks, rank, world)
    per_rank = (n_blocks + world - 1) // world
    sizes = torch.zeros(per_rank, dtype=torch.int32, device=device)
    for k, b in enumerate(mine):
        sizes[k] = len(enc[b])
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    width = int(max(int(s.max()) for s in all_sizes))
    payload = torch.zeros((per_rank, max(width, 1)), dtype=torch.uint8, device=device)
    for k, b in enumerate(mine):
        payload[k, :len(enc[b])] = torch.from_numpy(np.frombuffer(enc[b], dtype=np.uint8).copy()).to(device)
    gathered = [torch.zeros_like(payload) for _ in range(world)] if rank == 0 else None
    dist.gather(payload, gathered, dst=0)
    if rank != 0:
        return None
    out = {}
    for r in range(world):
        rows = gathered[r].cpu().numpy()
        for k, b in enumerate(blocks_of_rank(n_blocks, r, world)):
            out[b] = rows[k, :int(all_sizes[r][k])].tobytes()
    return out

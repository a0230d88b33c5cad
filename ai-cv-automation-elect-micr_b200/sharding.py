"""Image-wise sharding of independent micrographs / crops across the GPUs of one box.

The denoiser path has no exchange step: every crop's forward pass depends only on that crop, and overlap
averaging happens inside one image.  So the multi-GPU form is one process per GPU (launched with
``torch.distributed.run``), each owning one ``Engine``; item k goes to rank ``k % world`` and the only
communication is a host-side gather of finished images (object gather over the process group -- gloo or NCCL --
never a data-path collective).  The reference has no inference data parallelism at all
(machine_learning/denoiser.py:591 pins one GPU); this is the B200-box extension of ``Denoiser.denoise`` to a
stream of micrographs.
"""
from __future__ import annotations

from typing import Callable, List, Sequence


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: item k belongs to rank k % world."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} of world {world}")
    return list(range(rank, n_items, world))


def run_sharded(items: Sequence, fn: Callable, rank: int, world: int, group=None, gather: bool = True, precomputed=None):
    """Apply ``fn`` to the items this rank owns; with ``gather`` every rank gets the full result list in item order.

    ``fn(item)`` is e.g. ``Denoiser.denoise``.  ``group`` is a torch.distributed process group (None = default).
    ``precomputed``: {item index: result} for this rank's items when the caller has already produced them in one batch."""
    mine = shard_indices(len(items), rank, world)
    local = [(k, precomputed[k] if precomputed is not None else fn(items[k])) for k in mine]
    if not gather or world == 1:
        out = [None] * len(items)
        for k, r in local:
            out[k] = r
        return out
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, local, group=group)
    out = [None] * len(items)
    for part in parts:
        for k, r in part:
            out[k] = r
    return out


def denoise_stream(denoiser, images: Sequence, rank: int = 0, world: int = 1, group=None, gather: bool = True, **kw):
    """BASELINE.json config 4: a stream of micrographs, image k on GPU k % world; each rank tiles, infers and stitches
    its own images end to end, results gathered on the host.  A rank's images go through ``Denoiser.denoise_many`` (copies
    of neighbouring images overlapped with the network passes) when they share one shape, else one ``Denoiser.denoise``
    call each -- the results are bit-identical either way, and identical to a single-GPU run."""
    mine = shard_indices(len(images), rank, world)
    shapes = {tuple(getattr(images[k], "shape", ())) for k in mine}
    if len(mine) > 1 and len(shapes) == 1 and hasattr(denoiser, "denoise_many"):
        results = denoiser.denoise_many([images[k] for k in mine], **kw)
        by_index = dict(zip(mine, results))
        return run_sharded(images, None, rank, world, group, gather, precomputed=by_index)
    return run_sharded(images, lambda img: denoiser.denoise(img, **kw), rank, world, group, gather)

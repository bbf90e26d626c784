"""Multi-GPU plumbing: one process per GPU (torch.distributed), the frame sharded by SAMPLES, and exactly
one exchange step per frame.

Every (pixel, sample) pair owns a Philox counter, so the union of the ranks' sample subsets is bit-for-bit
the single-GPU sample set.  Rank r renders the strata rows s_j with s_j % world == r of
every pixel into a full-frame float4 accumulation buffer; the buffers are combined once per frame:

  combine="reduce"  one SUM reduce to rank 0 (NCCL over NVLink on GPUs, gloo on CPU tests)
  combine="gather"  all ranks' buffers gathered on rank 0 and summed in rank order — the same bits for any
                    world size that divides the strata rows evenly, at world x the traffic

With `exact_accum=1` the partial frames are (H, W, 4) int64 tensors of exact fixed-point sums: integer addition is
associative, so either way of combining gives the single-GPU frame bit for bit (Renderer.resolve_exact_device turns the
combined buffer into the float4 accumulation image).  The renderer has no other collective.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def sample_split(rank: int, world: int):
    """(sample_mod, sample_rem) for mort_render_opts."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return world, rank


def tile_split(rank: int, world: int):
    """(tile_mod, tile_rem) for mort_render_opts: rank r renders the 8-row bands b with b % world == r (round robin, so sky
    and geometry are spread over all ranks) into an otherwise untouched — caller-zeroed — full-frame buffer; the same SUM
    reduce (or a gather of the bands) assembles the frame.  Every pixel comes from exactly one rank: the assembled frame is
    bit-identical to the single-GPU frame even for the float accumulation image."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return world, rank


def rows_of_rank(sqrt_spp: int, rank: int, world: int) -> int:
    return len(range(rank, sqrt_spp, world))


def combine(accum: torch.Tensor, group=None, how: str = "reduce", dst: int = 0) -> torch.Tensor | None:
    """Combine per-rank partial accumulation buffers (H, W, 4) on `dst`.  Returns the combined tensor on
    dst and None elsewhere.  With one rank (or no process group) it is the identity."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return accum
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if how == "reduce":
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
        return accum if rank == dst else None
    if how == "gather":
        parts = [torch.empty_like(accum) for _ in range(world)] if rank == dst else None
        dist.gather(accum, parts, dst=dst, group=group)
        if rank != dst:
            return None
        total = parts[0].clone()
        for p in parts[1:]:
            total += p                      # fixed rank order
        return total
    raise ValueError(how)


def combine_virtual(partials: list[torch.Tensor]) -> torch.Tensor:
    """The same fixed-order sum for N "virtual ranks" executed serially on one device (single-GPU tests)."""
    total = partials[0].clone()
    for p in partials[1:]:
        total += p
    return total

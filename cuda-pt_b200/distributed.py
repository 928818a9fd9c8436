"""Multi-GPU plumbing: one process per GPU (torchrun), one NCCL collective on the HDR accumulator.

The reference's multi-GPU driver (scripts/ddp_render.py:70-81,141-143,161-166) gives every rank the FULL
frame with a different `seed_offset` and all-reduces `image * spp` plus the scalar spp.  Two schemes are
provided here, both without any collective on the per-bounce path (samples are independent):

* `sample_range(total_spp, world, rank)` — deterministic sample split: rank r renders sample numbers
  [base+1 .. base+n] of the SAME seed stream (seed = pixel + seed_offset + sample * 11467,
  wavefront_pt.cu:82), so the union over ranks is exactly the set of paths one GPU would trace; the reduced
  image equals the single-GPU image up to fp32 summation order.
* `reference_seed_offset(rank, base)` — the reference's own scheme (`local_rank * 4201 + base`).

`reduce_accumulator` sums the per-rank SUM images (`DeviceImage`, not the running mean) in place with a single
all-reduce; `torch.distributed` is used for the collective only (backend nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def sample_range(total_spp, world, rank):
    """Contiguous, balanced split of sample numbers 1..total_spp.  Returns (counter_base, count)."""
    if world < 1 or not (0 <= rank < world) or total_spp < 0:
        raise ValueError("bad sample split")
    q, r = divmod(int(total_spp), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def row_range(height, world, rank, align=4):
    """Tile split (SURVEY.md 8e "tile partitioning"): contiguous bands of rows, multiples of `align` rows (the ray
    generation kernel works on 8x4-pixel tiles), balanced to within one band unit.  Returns (row_begin, row_end)."""
    if world < 1 or not (0 <= rank < world) or height <= 0:
        raise ValueError("bad row split")
    units = (int(height) + align - 1) // align
    q, r = divmod(units, int(world))
    lo = rank * q + min(rank, r)
    hi = lo + q + (1 if rank < r else 0)
    return min(lo * align, height), min(hi * align, height)


def reference_seed_offset(rank, base_offset=0):
    """ddp_render.py:141-143 — `local_rank * 4201 + <random base>` (the base is passed in, not drawn here)."""
    return int(rank) * 4201 + int(base_offset)


def reduce_sum_(tensor, group=None, dst=None):
    """In-place SUM of a per-rank accumulator tensor (all ranks, or to `dst` only)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensor
    if dst is None:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def reduce_rendered_image(image, spp, group=None):
    """Reference semantics (ddp_render.py:70-81): all-reduce image*spp and spp, return the weighted mean."""
    acc = image * float(spp)
    n = torch.tensor([float(spp)], dtype=torch.float32, device=image.device)
    reduce_sum_(acc, group)
    reduce_sum_(n, group)
    return acc / n, int(n.item())


def render_distributed(renderer, total_spp, gamma=False, group=None, split="samples"):
    """Render `total_spp` samples split over the ranks of `group` and return the reduced mean image
    (torch tensor [H, W, 4] on the renderer's device, identical on every rank).

    split="samples": every rank renders the full frame for its share of the sample numbers (perfectly balanced; equals
    the 1-GPU image up to fp32 summation order).  split="tiles": every rank renders all samples of its band of rows
    (disjoint pixels, so the zero-filled SUM reproduces the 1-GPU image bit for bit; load follows the image content).
    `renderer` is a cuda_pt_b200.WavefrontRenderer freshly created or reset (counter 0)."""
    from .renderer import device_tensor
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if split == "tiles":
        lo, hi = row_range(renderer.height, world, rank)
        base, count = 0, int(total_spp) if hi > lo else 0
        if hi > lo:
            renderer.set_option("row_end", hi)
            renderer.set_option("row_begin", lo)
    elif split == "samples":
        base, count = sample_range(total_spp, world, rank)
    else:
        raise ValueError("split must be 'samples' or 'tiles'")
    renderer.set_counter(base)  # this rank's sample numbers are base+1 .. base+count
    if count > 0:
        renderer.render_batch(count, False)
    acc = device_tensor(renderer.accumulator_ptr(), (renderer.height, renderer.width, 4), renderer.device, renderer)
    total = acc.clone()
    torch.cuda.current_stream(total.device).synchronize()
    reduce_sum_(total, group)
    mean = total / float(total_spp)
    if gamma:
        mean = torch.cat([mean[..., :3].clamp_min(0).pow(1.0 / 2.2), torch.ones_like(mean[..., 3:])], dim=-1)
    return mean

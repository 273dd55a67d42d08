"""Multi-GPU layout of the try-on path: embarrassingly data-parallel.  One process per GPU, a full weight
replica each, the batch split contiguously; the ONLY collective is the gather of the finished images.  The reference's own
multi-process mode (accelerate, /root/reference/src/inference.py:223) has no collective at all and reseeds every rank
identically; here the three noise tensors are drawn ONCE for the full batch in the reference's order and sliced, so an
N-GPU run sees exactly the per-sample noise of the 1-GPU run.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous split; the first `total % world` ranks get one extra sample."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_inputs(inputs, rank, world):
    lo, hi = shard_bounds(next(iter(inputs.values())).shape[0], rank, world)
    return {k: v[lo:hi] for k, v in inputs.items()}


def draw_noise(total, h, w, generator, device="cpu"):
    """The pipeline's three RNG draws for the FULL batch, in the reference's order (tryon_pipe.py:640, 419, 458):
    cloth posterior, initial latents, masked-image posterior."""
    gdev = generator.device if generator is not None else torch.device(device)
    return tuple(torch.randn((total, 4, h, w), generator=generator, device=gdev, dtype=torch.float32) for _ in range(3))


def shard_noise(noise, rank, world):
    lo, hi = shard_bounds(noise[0].shape[0], rank, world)
    return tuple(n[lo:hi] for n in noise)


def gather_images(local, world, out=None):
    """All-gather of the per-rank image tensors [b_r, H, W, 3] (equal b_r: one NCCL all_gather_into_tensor over NVLink;
    ragged: all_gather of a padded tensor).  Returns the full batch on every rank."""
    if world == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1 and dist.get_backend() == "nccl":
        out = out if out is not None else torch.empty((sum(sizes),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])

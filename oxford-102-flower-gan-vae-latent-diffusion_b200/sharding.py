"""Multi-GPU sampling: the batch shards over ranks with NO per-step communication (every op of the v2
path is row-wise over the batch); only the decoded images are gathered, with one
all_gather_into_tensor (NCCL over NVLink on GPUs; gloo in the CPU tests of the host logic).

Noise is keyed by the GLOBAL sample index (oracle/philox.py), so the result for sample i is the same
for any number of ranks."""
import torch
import torch.distributed as dist


def shard_bounds(total, world_size, rank):
    """Contiguous, balanced [lo, hi) of `total` samples for `rank`: the first total % world_size ranks get
    one extra sample."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside [0, %d)" % (rank, world_size))
    base, extra = divmod(total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local, total, group=None):
    """All-gather row shards produced with shard_bounds() into the full (total, ...) tensor on every rank.
    Uses ONE all_gather_into_tensor; ragged shards are padded to the largest shard and trimmed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    per = -(-total // ws)
    pad = per - local.shape[0]
    if pad:
        local = torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((ws * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if total % ws == 0:
        return out
    pieces = []
    for r in range(ws):
        lo, hi = shard_bounds(total, ws, r)
        pieces.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)


def generate_sharded(autoencoder, diffusion, classes, *, seed, device=None, group=None, gather=True):
    """Sample + decode `classes` (global int64 tensor of length B, identical on every rank): rank r
    denoises rows shard_bounds(B, world, r) with global Philox indices, decodes them, and (gather=True)
    all-gathers the images.  Returns (images, local_latents, (lo, hi))."""
    B = int(classes.shape[0])
    ws = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rk = dist.get_rank(group) if ws > 1 else 0
    lo, hi = shard_bounds(B, ws, rk)
    device = device if device is not None else next(autoencoder.parameters()).device
    c_local = classes[lo:hi].to(device)
    with torch.no_grad():
        latents = diffusion.sample((hi - lo, autoencoder.latent_dim), device, c_local, seed=seed, sample_offset=lo)
        images = autoencoder.decode(latents)
    if gather:
        images = gather_rows(images, B, group)
    return images, latents, (lo, hi)

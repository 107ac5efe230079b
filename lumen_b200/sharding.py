"""Sample-index sharding of a render across GPUs (one process per GPU) -- SURVEY.md section 8e.

Every (pixel, frame) sample is independent: the RNG seed is (x, y, frame, 0) (path.rgen:23) and the film update is
per-pixel, so rank r of N renders frames {first + r + k*N}. Each rank holds a full scene + BVH replica and accumulates an
un-normalised SUM with a per-pixel valid-sample count in alpha (the reference skips NaN samples, path.rgen:102-104). The
only exchange step is one all-reduce (sum, fp32) of the W*H*4 film, after which every rank divides rgb by the count.
No other data-path collective exists.

Secondary axis, pixel tiles (BASELINE config 4, "tile + sample sharding"): the N ranks form a grid of `pixel_shards` x
`sample_shards`; pixel shard p renders image rows p, p + P, p + 2P, ... (lmb_set_pixel_shard: full-width tiles one row high,
interleaved so every shard sees the same mix of content) and, inside it, sample shard s takes frames first + s + k*S. Every
rank still writes into a full-size sum film whose foreign rows stay zero, so the exchange step is the same single all-reduce.
"""
import numpy as np


def frames_of_rank(first_frame, frames_per_rank, rank, world):
    """Frame indices rank `rank` renders: (first, stride, count) -> first + rank, first + rank + world, ..."""
    return first_frame + rank, world, frames_per_rank


def shard_frame_list(first_frame, frames_per_rank, rank, world):
    f0, stride, n = frames_of_rank(first_frame, frames_per_rank, rank, world)
    return [f0 + k * stride for k in range(n)]


def grid_of_rank(rank, world, pixel_shards):
    """(pixel shard p, sample shard s, P, S) of `rank` in a P x S grid, P * S == world. Ranks that share a pixel shard are
    consecutive, so they differ only in the frames they render."""
    if pixel_shards < 1 or world % pixel_shards:
        raise ValueError(f"pixel_shards={pixel_shards} does not divide world={world}")
    sample_shards = world // pixel_shards
    return rank // sample_shards, rank % sample_shards, pixel_shards, sample_shards


def rows_of_shard(height, row_first, row_stride):
    """Image rows a pixel shard owns."""
    return np.arange(row_first, height, row_stride)


def accumulate_sum(film_rgba, sample_rgb):
    """Host-side statement of LMB_FILM_SUM for one frame: rgb += sample, alpha += 1 unless luminance is NaN."""
    lum = sample_rgb[..., 0] * np.float32(0.2126) + sample_rgb[..., 1] * np.float32(0.7152) + sample_rgb[..., 2] * np.float32(0.0722)
    ok = ~np.isnan(lum)
    film_rgba[..., :3][ok] += sample_rgb[ok]
    film_rgba[..., 3][ok] += 1.0
    return film_rgba


def resolve(film_rgba):
    """lmb_resolve: rgb /= alpha where alpha > 0, alpha = 1."""
    out = np.zeros_like(film_rgba)
    a = film_rgba[..., 3]
    nz = a > 0
    out[..., :3][nz] = film_rgba[..., :3][nz] / a[nz][:, None]
    out[..., 3] = 1.0
    return out


def all_reduce_film(film_tensor, dist):
    """The single collective of the path: in-place fp32 sum of the film over all ranks (NCCL on GPUs, gloo in CPU tests)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(film_tensor, op=dist.ReduceOp.SUM)
    return film_tensor

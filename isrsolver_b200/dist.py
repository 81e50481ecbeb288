"""One-process-per-GPU sharding of toy campaigns (torch.distributed is plumbing only).

The path shards with no data-path collective: every rank assembles and factors the same N x N problem
(replicated: it is << 1 ms of GPU work) and processes a contiguous slice of the toys (or of the lambda
grid).  The single exchange step is ONE all_gather of a packed per-rank buffer
    [ chi2 shard (padded to the largest shard) | sum_bcs (N) | ratio {count, sum, sum2} (3N) ]
after which every rank holds all chi2 values and computes the TH1F(128, min, max) histogram exactly as
src/Chi2Test.cpp:64-78 does on one process -- the histogram range depends on the GLOBAL min/max, so
binning must happen after the gather to stay bit-identical with the single-GPU result.

When only the HISTOGRAM is wanted (the chi2 test itself), the chi2 values need not move at all
(`histogram_sharded`): one all_gather of the packed per-rank summary [min, max | sum_bcs | ratio stats], local
TH1F counts against the global range, one all_reduce of the integer counters -- what bench.py does per step.
Device-drawn campaigns (isr_toy_campaign) shard by `first_toy`; lambda grids shard by slicing the grid.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous [begin, end) slice of `total` units for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pack(chi2, sum_bcs, ratio_stats, max_shard: int):
    """Pack one rank's results into a flat float64 buffer of fixed length."""
    n = len(sum_bcs)
    buf = np.zeros(max_shard + 4 * n)
    buf[: len(chi2)] = chi2
    buf[max_shard: max_shard + n] = sum_bcs
    buf[max_shard + n:] = np.asarray(ratio_stats).reshape(-1)
    return buf


def unpack(gathered, total: int, world: int, n: int):
    """Inverse of pack over the gathered (world, len) array: all chi2 in toy order + summed statistics."""
    max_shard = gathered.shape[1] - 4 * n
    parts, sum_bcs, stats = [], np.zeros(n), np.zeros(3 * n)
    for r in range(world):
        b, e = shard_bounds(total, world, r)
        parts.append(gathered[r, : e - b])
        sum_bcs += gathered[r, max_shard: max_shard + n]      # fixed rank order => deterministic
        stats += gathered[r, max_shard + n:]
    return np.concatenate(parts), sum_bcs, stats.reshape(n, 3)


def all_gather_results(chi2, sum_bcs, ratio_stats, total: int, device=None):
    """The one collective of a toy campaign.  Works on NCCL (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    n = len(sum_bcs)
    max_shard = max(shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world))
    mine = torch.from_numpy(pack(chi2, sum_bcs, ratio_stats, max_shard))
    if device is not None:
        mine = mine.to(device)
    out = torch.empty(world * mine.numel(), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(out, mine)       # the single collective
    return unpack(out.view(world, mine.numel()).cpu().numpy(), total, world, n)


def histogram_sharded(chi2_local, sum_bcs, ratio_stats, nbins: int = 128, count_fn=None, device=None, local_range=None):
    """The chi2 test over all ranks' toys without gathering them.

    Returns (counts[nbins + 2] int64 summed over ranks, lo, hi, sum_bcs, ratio_stats): bit-identical to the
    single-process TH1F(nbins, min, max) of the concatenated values because every rank bins against the GLOBAL
    range and integer counters add exactly (sum_bcs / ratio_stats are float sums re-associated across ranks: equal
    to rounding, not bit for bit).  count_fn(lo, hi) -> counts lets the caller bin on the GPU (isr_histogram_dev);
    the default bins `chi2_local` on the host.  local_range = (lo, hi): this rank's range when the chi2 values stay
    on the device (chi2_local may then be empty).  An EMPTY shard (fewer toys than ranks) contributes (+inf, -inf)
    and zero counts.  Two collectives, both O(N) bytes."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    chi2_local = np.asarray(chi2_local, dtype=np.float64)
    n = len(sum_bcs)
    if local_range is not None:
        lo, hi = float(local_range[0]), float(local_range[1])
    else:
        lo = chi2_local.min() if chi2_local.size else np.inf
        hi = chi2_local.max() if chi2_local.size else -np.inf
    mine = torch.from_numpy(np.concatenate([[lo, hi], np.asarray(sum_bcs, dtype=np.float64),
                                            np.asarray(ratio_stats, dtype=np.float64).reshape(-1)]))
    if device is not None:
        mine = mine.to(device)
    allr = torch.empty(world * mine.numel(), dtype=torch.float64, device=mine.device)
    dist.all_gather_into_tensor(allr, mine)                       # collective 1: per-rank summaries
    a = allr.view(world, -1).cpu().numpy()
    glo, ghi = float(a[:, 0].min()), float(a[:, 1].max())
    sb = a[:, 2:2 + n].sum(0)                                     # fixed rank order => deterministic
    rs = a[:, 2 + n:].sum(0).reshape(n, 3)
    counts = count_fn(glo, ghi) if count_fn else th1_counts(chi2_local, nbins, glo, ghi)
    c = torch.from_numpy(np.asarray(counts).astype(np.int64))
    if device is not None:
        c = c.to(device)
    dist.all_reduce(c)                                            # collective 2: integer counters
    return c.cpu().numpy(), glo, ghi, sb, rs


def th1_counts(values, nbins: int, lo: float, hi: float):
    """TH1F::Fill bin arithmetic on the host (used by the gloo tests and as the cross-check of the device
    histogram): counts[nbins + 2] with under/overflow."""
    v = np.asarray(values, dtype=np.float64)
    idx = np.empty(v.shape, dtype=np.int64)
    below, above = v < lo, ~(v < hi)
    mid = ~(below | above)
    idx[below], idx[above] = 0, nbins + 1
    idx[mid] = 1 + (nbins * (v[mid] - lo) / (hi - lo)).astype(np.int64)
    return np.bincount(idx, minlength=nbins + 2).astype(np.uint32)

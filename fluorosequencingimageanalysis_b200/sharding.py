"""
Host-side multi-GPU partitioner (replaces the process fan-out of pflib.parallel_image_batch,
pflib.py:1000-1111).  Every (field, cycle/frame) image is independent for detection, fitting
and metrics (pflib.image_batch treats files independently, pflib.py:957-996), so the path
shards with NO collective: one process per GPU, contiguous blocks of fields per rank (all
cycles of a field stay on one GPU so the later tracking/photometry rows need no exchange),
results concatenated on the host.  torch.distributed is used only for the rendezvous,
barriers and the final gather of small packed results.
"""
import numpy as np


def balance_by_count(counts, n_parts):
    """The greedy balancing of pflib.py:1056-1069: images sorted by DEcreasing candidate count,
    popped from the END of that list (i.e. smallest first -- the reference's quirk), each
    appended to the partition whose candidate total is currently smallest (first such
    partition on ties, as `sorted(...)[0]` is stable).  Returns a list of index lists."""
    if n_parts < 1 or round(n_parts) != n_parts:
        raise ValueError("Number of processes must be an integer >= 1")      # pflib.py:1059-1060
    order = sorted(range(len(counts)), key=lambda i: counts[i], reverse=True)
    parts = [[] for _ in range(int(n_parts))]
    sums = [0] * int(n_parts)
    while order:
        i = order.pop()
        k = min(range(len(parts)), key=lambda p: sums[p])
        parts[k].append(i)
        sums[k] += counts[i]
    return parts


def field_block(n_fields, world_size, rank):
    """Contiguous block [lo, hi) of fields owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_fields, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_field_blocks(cand_counts_per_field, world_size):
    """Contiguous field blocks with near-equal candidate totals (prefix-sum split); keeps the
    all-cycles-of-a-field-on-one-GPU property while balancing fit work like
    parallel_image_batch does.  Returns [(lo, hi)] per rank."""
    c = np.asarray(cand_counts_per_field, dtype=np.float64)
    n = len(c)
    cs = np.concatenate([[0.0], np.cumsum(c)])
    total = cs[-1]
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(cs, target, side="left"))
        if k > 0 and k <= n and (target - cs[k - 1]) <= (cs[min(k, n)] - target):
            k -= 1                                    # nearer prefix sum
        k = min(max(k, bounds[-1]), n)
        bounds.append(k)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def gather_packed(local_arrays, group=None):
    """Concatenate per-rank packed numpy arrays on every rank (rank order).  Uses
    torch.distributed.all_gather_object on the default (NCCL or gloo) group: results are
    gathered to the host only at the end; no data-path collective exists."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_arrays
    ws = dist.get_world_size(group)
    bucket = [None] * ws
    dist.all_gather_object(bucket, local_arrays, group=group)
    out = {}
    for k in local_arrays:
        out[k] = np.concatenate([b[k] for b in bucket], axis=0)
    return out

"""Multi-GPU sharding of the hot path: one process per GPU, NO data-path collective.

The reference's batch loop (hw2.cpp:328-338) carries no state from one pair to the next and hw3's distance
stage (hw3.cpp:231-241) couples its 120 pairs only through 16 sums added up afterwards, so both shard by
contiguous index ranges.  What crosses ranks is tiny and host-side: the per-rank winner (hw2.cpp:340-357) or the
per-rank partial star sums (hw3.cpp:238-239).  torch.distributed carries it (backend nccl on the GPU box, gloo in
the CPU tests); ties are resolved exactly as the serial reference does -- lowest pair index wins.
"""
import numpy as np
import torch
import torch.distributed as dist

GLOBAL, LOCAL = 0, 1


def pair_range(n_pairs, rank, world):
    """Contiguous shard [first, first + count) of the pair index space (same split as csrc/hw2_main.cpp)."""
    first = n_pairs * rank // world
    return first, n_pairs * (rank + 1) // world - first


def star_pair_range(n_seqs, rank, world):
    """Shard of the i < j pair list of hw3.cpp:233-234 (row-major order, as b2a_affine_star_scores numbers it)."""
    return pair_range(n_seqs * (n_seqs - 1) // 2, rank, world)


def local_best(mode, results):
    """(key, local index) of a shard's winner: overlap (global) / score (local), strict '>' from -1000000 (hw2.cpp:326-357)."""
    key = results["overlap"] if mode == GLOBAL else results["score"]
    best, idx = -1000000, -1
    if len(key):
        k = int(np.argmax(key))                 # numpy returns the first maximum
        if int(key[k]) > best:
            best, idx = int(key[k]), k
    return best, idx


def first_strict_max(candidates):
    """Batch winner from per-shard winners: candidates[r] = (key, global pair index or -1) of shard r, shards own ascending index ranges.
    hw2.cpp:326-357 scans all pairs with a strict '>' from -1000000, so the winner is the first strict maximum of the candidates in shard
    order (ties resolve to the lowest index).  Returns the global index, -1 if no pair beats -1000000."""
    best_key, best_idx = -1000000, -1
    for key, idx in candidates:
        if int(idx) >= 0 and int(key) > best_key:
            best_key, best_idx = int(key), int(idx)
    return best_idx


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def merge_best(mode, results, first):
    """Global winner over all ranks' shards: returns (global pair index or -1, key).  Every rank gets the answer.
    Ranks own ascending, contiguous index ranges, so 'first strict maximum' = highest key, then lowest index."""
    key, idx = local_best(mode, results)
    gidx = first + idx if idx >= 0 else -1
    mine = torch.tensor([key, gidx], dtype=torch.int64, device=_device())
    if dist.is_initialized() and dist.get_world_size() > 1:
        allv = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    best_key, best_idx = -1000000, -1
    for v in allv:
        k, g = int(v[0]), int(v[1])
        if g >= 0 and (k > best_key or (k == best_key and best_idx >= 0 and g < best_idx)):
            best_key, best_idx = k, g
    return best_idx, best_key


def reduce_star_sums(partial_sums):
    """Sum of the per-rank partial star sums (hw3.cpp:238-239) and the centre index (hw3.cpp:243-251)."""
    t = torch.as_tensor(np.asarray(partial_sums, dtype=np.int64), device=_device())
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    sums = t.cpu().numpy().astype(np.int32)       # the reference adds in int
    centre = 0
    for i in range(1, len(sums)):
        if sums[i] > sums[centre]:
            centre = i
    return sums, (centre if len(sums) else -1)


def max_over_ranks(x):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())

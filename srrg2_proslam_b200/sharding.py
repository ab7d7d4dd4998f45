"""Multi-GPU partitioning of the frontend path: one process per GPU, `torch.distributed` for the plumbing.

* detect / describe / stereo match: frames are independent units -> contiguous frame ranges per rank, NO data-path
  collective; variable-length results are gathered host-side.
* brute-force Hamming sweep: query rows sharded, train set replicated; each rank emits per-row (best, second, argmin)
  and ONE all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests) assembles the full table.
* projective match + linearise + H,b: sequential per sequence -> replicas only (nothing here).
"""
import torch
import torch.distributed as dist


def frame_range(n_frames, rank, world):
    """contiguous frame range [begin, end) of `rank`; the first n_frames % world ranks own one extra frame"""
    q, r = divmod(int(n_frames), int(world))
    begin = rank * q + min(rank, r)
    return begin, begin + q + (1 if rank < r else 0)


def query_rows(n_query, rank, world, align=256):
    """query-row shard [begin, end) of `rank`: equal shards of ceil(n_query / world) rows rounded up to `align`
    (the sweep kernel handles 256 queries per CTA); trailing ranks may own fewer (or zero) rows"""
    per = -(-int(n_query) // int(world))
    per = -(-per // align) * align
    begin = min(rank * per, n_query)
    return begin, min(begin + per, n_query)


def allgather_best2(local, n_query, align=256, group=None):
    """local: int32 [3, rows_local] = (best, second, argmin) of this rank's query rows (device or host tensor).
    Returns int32 [3, n_query] on every rank.  One all_gather_into_tensor of equal, padded shards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local[:, :n_query]
    per = query_rows(n_query, 0, world, align)[1]
    pad = torch.zeros((3, per), dtype=torch.int32, device=local.device)
    pad[:, :local.shape[1]] = local
    out = torch.empty((world * 3, per), dtype=torch.int32, device=local.device)  # rank-major concatenation along dim 0
    dist.all_gather_into_tensor(out, pad, group=group)
    return out.view(world, 3, per).permute(1, 0, 2).reshape(3, world * per)[:, :n_query].contiguous()


def gather_frame_results(local_counts, local_points, group=None):
    """host-side gather of variable-length per-frame results: every rank contributes (counts[int64 n_local],
    points[float32 n_points, k]) for its frame range; rank 0 receives them concatenated in frame order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_counts, local_points
    rank = dist.get_rank(group)
    objs = [None] * world if rank == 0 else None
    dist.gather_object((local_counts.cpu(), local_points.cpu()), objs, dst=0, group=group)
    if rank != 0:
        return None, None
    return torch.cat([o[0] for o in objs]), torch.cat([o[1] for o in objs])

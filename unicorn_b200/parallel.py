"""Multi-GPU execution: one process per GPU, sequences sharded across ranks, no per-frame collective.

The reference's only inference-time parallelism is per-video sharding (`DistributedVideoSampler`,
external/qdtrack/qdtrack/datasets/samplers/distributed_video_sampler.py:13-25; the SOT pool in
external/lib/test/evaluation/running.py:111-118,199-202).  A frame is never split across GPUs (whole-frame GroupNorm
statistics, 2.5 ms budget), so the fabric is used only for a start barrier and one fixed-size gather of per-rank
results/timings at the end (replacing the pickled gloo gathers of unicorn/utils/dist.py:224-265 and
dist.reduce(statistics) of unicorn/evaluators/mot_evaluator.py:241).
"""
import torch
import torch.distributed as dist


def shard_sequences(n_sequences, rank, world_size):
    """Contiguous-chunk partition like DistributedVideoSampler: rank r owns chunk r of ceil-sized chunks; every
    sequence is owned by exactly one rank."""
    base, rem = divmod(n_sequences, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def gather_stats(frames, seconds, n_tracks, device=None):
    """All ranks contribute [frames, seconds, n_tracks]; returns (total_frames, max_seconds, total_tracks, per_rank).
    One all_gather of 3 doubles per rank (NCCL over NVLink on GPUs, gloo on CPU)."""
    t = torch.tensor([float(frames), float(seconds), float(n_tracks)], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()):
        return frames, seconds, n_tracks, [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    per_rank = [o.tolist() for o in out]
    return (sum(p[0] for p in per_rank), max(p[1] for p in per_rank), sum(p[2] for p in per_rank), per_rank)


def run_sharded(sequences, worker, device=None):
    """Run `worker(seq_index, sequence) -> (frames, n_tracks)` over this rank's shard between two barriers and
    return aggregate frames/s computed from the slowest rank (device timing is the caller's business)."""
    import time
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = shard_sequences(len(sequences), rank, world)
    if dist.is_initialized():
        dist.barrier()
    t0 = time.perf_counter()
    frames = tracks = 0
    for i in mine:
        f, k = worker(i, sequences[i])
        frames += f
        tracks += k
    dt = time.perf_counter() - t0
    tot_f, max_t, tot_k, per_rank = gather_stats(frames, dt, tracks, device)
    return dict(frames=tot_f, seconds=max_t, tracks=tot_k, fps=tot_f / max(max_t, 1e-9), per_rank=per_rank, shard=mine)

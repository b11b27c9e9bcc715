"""Env-parallel data parallelism: one process per GPU, environments sharded across ranks, the
only data-path collective is the all-reduce of the flat PPO gradient (NCCL over NVLink/NVSwitch on the GPU box; the
same host logic runs on gloo in the CPU tests).  The reference has no distributed code at all -- this is the first
real collective of the path."""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_envs(total_envs: int, rank: int, world_size: int):
    """Contiguous shard [start, start+count) of `total_envs` for `rank`; shards differ by at most one env."""
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def allreduce_mean_(flat: torch.Tensor):
    """In-place mean over ranks of one flat bucket (the whole 4.6 MB / 59 MB gradient arena in a single call: on
    NVSwitch the cost is launch latency, not link count)."""
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / ws)
    return flat


def allreduce_avg_(t: torch.Tensor):
    """In-place average over ranks, as one collective: NCCL reduces with AVG; gloo (CPU tests) has no AVG, so SUM and
    scale.  Enqueued on the CURRENT stream's NCCL work queue (call it under `torch.cuda.stream(side)` to overlap)."""
    rank, ws = world()
    if ws > 1:
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / ws)
    return t


def require_equal(value: int, what: str):
    """Raises on every rank if `value` differs across ranks (e.g. rollout sizes: unequal minibatch counts would leave the
    per-minibatch gradient all-reduces unmatched and hang the job)."""
    rank, ws = world()
    if ws > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        lo = torch.tensor([int(value)], dtype=torch.int64, device=dev)
        hi = lo.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if int(lo) != int(hi):
            raise RuntimeError(f"{what} differs across ranks ({int(lo)} .. {int(hi)}): every rank must run the same number of "
                               "minibatches per epoch; use equal env shards")


def consensus_max(x: torch.Tensor):
    """MAX over ranks -- used for approx_kl so that the KL early stop of ppo_grid_obs.py:264-268 is taken by every rank
    at the same minibatch (a rank-local `break` would dead-lock the next gradient all-reduce)."""
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return x


def should_stop(approx_kl: torch.Tensor, target_kl):
    """True on every rank iff any rank's approx_kl exceeds 1.5 * target_kl."""
    if target_kl is None:
        return False
    kl = consensus_max(approx_kl.detach().clone().reshape(1))
    return bool(float(kl) > 1.5 * target_kl)


def broadcast_state_(flat_params: torch.Tensor, buffers):
    """Start every rank from rank 0's parameters and BatchNorm buffers."""
    rank, ws = world()
    if ws > 1:
        dist.broadcast(flat_params, 0)
        for b in buffers:
            dist.broadcast(b, 0)


class OverlappedGradAllreduce:
    """The per-minibatch gradient all-reduce (mean) of the flat bucket `[vote slot | conv tensors | Linear tensors]`
    (gennbv_b200/policy.py) in two pieces: `start_linear()` right after the Linear phase of the backward -- it runs on a
    side stream under the convolution backward and carries 99.9 % of the bytes -- and `finish()` after the conv phase: the
    remaining ~30 KB (votes + conv tensors) on the caller's stream, then the join."""

    def __init__(self, bucket: torch.Tensor, split: int):
        self.big, self.small = bucket[split:], bucket[:split]
        self.stream = None

    def start_linear(self):
        if world()[1] == 1:
            return
        main = torch.cuda.current_stream()
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        self.stream.wait_stream(main)
        with torch.cuda.stream(self.stream):
            allreduce_avg_(self.big)

    def finish(self):
        if world()[1] == 1:
            return
        allreduce_avg_(self.small)
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

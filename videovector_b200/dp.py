"""Host-side data-parallel logic (pure numpy / torch.distributed, no CUDA): how a global batch is split over
ranks and how per-rank results combine.  The device side is vv_trainer_step's NCCL all-reduce.

Two ways to feed G ranks:
  * sharded streams (bench default): rank r owns videos [r*V/G, (r+1)*V/G) and runs its own reference-exact
    sampler over them -- no data-path collective, host sampling cost per rank stays constant;
  * global stream: one sampler stream of G*B items per step (identical to the 1-GPU / CPU index stream),
    rank r takes items [r*B, (r+1)*B).
Either way each rank normalises its loss by its LOCAL B*Nn (max_margin_loss_layer.cpp:117,191), so
sum-all-reduce followed by 1/G reproduces the global-batch mean."""
import numpy as np


def shard_videos(num_videos, rank, world):
    """[begin, end) of the videos rank `rank` owns."""
    per = num_videos // world
    return rank * per, (rank + 1) * per if rank < world - 1 else num_videos


def shard_batch(idx, quirk, rank, world):
    """Rank's slice of a global batch [G*B, R] produced by ONE sampler stream."""
    gb = idx.shape[0]
    assert gb % world == 0, "global batch must divide evenly over ranks"
    b = gb // world
    return idx[rank * b:(rank + 1) * b], quirk[rank * b:(rank + 1) * b]


def combine_gradients(local_grads, world):
    """What all-reduce(sum) + the 1/G scale inside vv_sgd_update compute, for host-side checks."""
    return sum(local_grads) / float(world)


def allreduce_mean_(tensors, world, dist):
    """In-place mean over ranks with torch.distributed (gloo on CPU in tests, NCCL on GPU)."""
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= world
    return tensors

"""Host-side data-parallel logic (pure numpy / torch.distributed, no CUDA): how a global batch is split over
ranks, which rows of W a rank owns in the peer-memory exchange, and how per-rank results combine.  The device side is
the exchange inside the update kernel (csrc/vv_dp_exchange.cu) or, in NCCL mode, vv_trainer_step's all-reduce.

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


def owner_rows(N, rank, world):
    """[begin, end) of the rows of W [N,K] (and of the momentum history) rank `rank` owns and updates in the peer-memory
    exchange (vv_dp_exchange.cuh): contiguous blocks of N/world rows; N must divide evenly."""
    assert N % world == 0, "N must divide evenly over the ranks"
    per = N // world
    return rank * per, (rank + 1) * per


def owner_update(contribs, W_rows, hist_rows, rate, momentum, decay, world):
    """What the owner does with the G contributions to its rows (phase B of the exchange): sum in rank order, scale by
    1/G, L2 decay, momentum, update.  numpy float32, same operation order as the kernel.  Returns (W_rows, hist_rows)."""
    g = contribs[0].astype(np.float32).copy()
    for c in contribs[1:]:
        g += c
    g *= np.float32(1.0 / world)
    g += np.float32(decay) * W_rows
    h = np.float32(rate) * g + np.float32(momentum) * hist_rows
    return W_rows - h, h


def combine_gradients(local_grads, world):
    """What all-reduce(sum) + the 1/G scale inside vv_sgd_update compute, for host-side checks."""
    return sum(local_grads) / float(world)


def allreduce_mean_(tensors, world, dist):
    """In-place mean over ranks with torch.distributed (gloo on CPU in tests, NCCL on GPU)."""
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= world
    return tensors

"""Data-parallel host logic with world_size 2 on the gloo backend (CPU): the per-rank local-mean gradients,
sum-all-reduced and scaled by 1/G, equal the global-batch gradient of the oracle, and the sharded sampler
streams stay inside their shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videovector_b200 import dp, ops


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, C, Nn, K, N = 6, 5, 10, 32, 16           # per-rank items
    R = C + Nn
    V, S = 40, 12
    vid, off, sid = ops.synthetic_videos(V, S)
    bank = ops.bank_host(V * S, K, 99)
    rng = np.random.RandomState(5)
    W = rng.normal(0, 0.1, (N, K)).astype(np.float32); b = rng.normal(0, 0.1, N).astype(np.float32)
    # ---- global stream: every rank runs the same sampler, takes its slice
    smp = ops.Sampler(vid, off, sid, B * world, C, Nn, 100, 50, 6, 100, rand_seed=1)
    gidx, gq = smp.next()
    idx, quirk = dp.shard_batch(gidx, gq, rank, world)
    gmask = (np.random.RandomState(9).uniform(0, 1, (R, B * world, N)) > 0.5).astype(np.uint32)   # [slot, item, n]
    mask = np.ascontiguousarray(gmask[:, rank * B:(rank + 1) * B]).reshape(R * B, N)

    def blob(ix, qk):
        g = bank[ix]
        g[..., K - 1] = np.where(qk >= 0, bank[np.maximum(qk, 0), K - 1], g[..., K - 1])
        g[..., K - 1] = np.where(qk == -1, 0.0, g[..., K - 1])
        return g
    loc = orc.net_forward_backward(blob(idx, quirk), W, b, mask, B, C, Nn, dropout_ratio=0.5, want=("loss", "violations", "dW", "db"))
    t = [torch.from_numpy(loc["dW"].copy()), torch.from_numpy(loc["db"].copy()), torch.from_numpy(loc["loss"].copy())]
    dp.allreduce_mean_(t, world, dist)
    viol = torch.from_numpy(loc["violations"].copy()); dist.all_reduce(viol)
    if rank == 0:
        glob = orc.net_forward_backward(blob(gidx, gq), W, b, gmask.reshape(R * B * world, N), B * world, C, Nn,
                                        dropout_ratio=0.5, want=("loss", "violations", "dW", "db"))
        q.put(dict(dW=float(np.abs(t[0].numpy() - glob["dW"]).max() / np.abs(glob["dW"]).max()),
                   db=float(np.abs(t[1].numpy() - glob["db"]).max() / np.abs(glob["db"]).max()),
                   loss=float(abs(t[2].item() - glob["loss"][0])), viol=float(viol.item() - glob["violations"][0])))
    # ---- the peer-memory exchange's protocol on the host (reduce-scatter by push, owner update, all-gather by push):
    # owner-sharded updates reproduce the replicated all-reduce + full update, and every rank ends with the same W
    r0, r1 = dp.owner_rows(N, rank, world)
    mine = torch.from_numpy(loc["dW"].copy())
    everyone = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(everyone, mine)                                    # "push": every rank's contribution, by source rank
    hist = np.zeros_like(W)
    Wr, hr = dp.owner_update([e.numpy()[r0:r1] for e in everyone], W[r0:r1], hist[r0:r1], 0.05, 0.9, 5e-4, world)
    rows = [torch.empty((N // world, K), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(rows, torch.from_numpy(np.ascontiguousarray(Wr)))  # "push" of the updated rows to every rank
    W_sharded = torch.cat(rows).numpy()
    W_repl, _, _ = orc.sgd_update(W, t[0].numpy(), hist, 0.05, 0.9, 5e-4)    # all-reduce mean + full update on one rank
    chk = torch.from_numpy(W_sharded.copy()); lo = chk.clone()
    dist.all_reduce(chk, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(dict(sharded_vs_replicated=float(np.abs(W_sharded - W_repl).max() / np.abs(W_repl).max()),
                   replicas_identical=bool(torch.equal(chk, lo))))
    # ---- sharded streams: each rank samples only from its own videos
    v0, v1 = dp.shard_videos(V, rank, world)
    s2 = ops.Sampler(vid[v0:v1], off[v0:v1 + 1] - off[v0], sid[off[v0]:off[v1]], B, C, Nn, 60, 50, 6, 100, rand_seed=1 + rank)
    i2, _ = s2.next()
    ok = torch.tensor([int(i2.min() >= 0 and i2.max() < off[v1] - off[v0])])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(dict(shard_ok=int(ok.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_two_ranks_gloo(vvlib, oracle):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(3):
        res.update(q.get(timeout=180))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res["dW"] < 2e-6 and res["db"] < 2e-6 and res["loss"] < 1e-6 and res["viol"] == 0.0, res
    assert res["sharded_vs_replicated"] < 1e-6 and res["replicas_identical"], res
    assert res["shard_ok"] == 1


def test_shard_helpers():
    assert dp.shard_videos(10, 0, 4) == (0, 2) and dp.shard_videos(10, 3, 4) == (6, 10)
    idx = np.arange(24).reshape(8, 3); q = -idx
    a, b = dp.shard_batch(idx, q, 1, 4)
    assert a.tolist() == [[6, 7, 8], [9, 10, 11]] and (b == -a).all()
    g = dp.combine_gradients([np.ones(3), 3 * np.ones(3)], 2)
    assert g.tolist() == [2, 2, 2]

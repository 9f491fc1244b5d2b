"""Run under torchrun with G ranks: G-rank data-parallel steps vs a 1-rank trainer on the global batch.
Prints 'DP_CHECK OK ...' on rank 0 (used by tests/test_gpu_dp.py and the multi-GPU GPU trip)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from videovector_b200 import ops, dp
from videovector_b200._lib import DROPOUT_MASK01

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
B, C, Nn, K, N = 64, 5, 10, 1024, 512          # per rank (N = 512: the sliced wgrad / all-reduce pipeline runs)
fused = prec in ("f16x3", "bf16") and os.environ.get("VV_DP_MATERIALISE") != "1"
R = C + Nn
V, S = 128, 24
vid, off, sid = ops.synthetic_videos(V, S)
bank = ops.fill_bank(V * S, K, 1234)
rng = np.random.RandomState(1701)
W0 = torch.as_tensor(rng.normal(0, 0.02, (N, K)).astype(np.float32)).cuda()
b0 = torch.as_tensor(rng.normal(0, 0.01, N).astype(np.float32)).cuda()
sol = dict(base_lr=0.05)
tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, prec=prec, dropout_ratio=0.5, dropout_mode=DROPOUT_MASK01,
                                 world_size=world, rank=rank, **sol))
tr.set_weights(W0, b0)
if fused:
    tr.set_bank(bank)
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(ops.dp_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
tr.dp_init(bytes(idt.cpu().numpy().tobytes()))
ref = None
if rank == 0:
    ref = ops.Trainer(ops.trainer_cfg(B * world, C, Nn, K, N, prec=prec, dropout_ratio=0.5, dropout_mode=DROPOUT_MASK01, **sol))
    ref.set_weights(W0, b0)
    if fused:
        ref.set_bank(bank)
smp = ops.Sampler(vid, off, sid, B * world, C, Nn, 500, 50, 6, 100, rand_seed=1)     # global stream, same on every rank
mrng = np.random.RandomState(3)
worst = 0.0
for it in range(4):
    gidx, gq = smp.next()
    gmask = (mrng.uniform(0, 1, (R, B * world, N)) > 0.5).astype(np.int32)
    idx, quirk = dp.shard_batch(gidx, gq, rank, world)
    mask = np.ascontiguousarray(gmask[:, rank * B:(rank + 1) * B]).reshape(R * B, N)
    tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), torch.as_tensor(mask).cuda(), it=it)
    if rank == 0:
        ref.step(bank, torch.as_tensor(gidx).cuda(), torch.as_tensor(gq).cuda(), torch.as_tensor(gmask.reshape(R * B * world, N)).cuda(), it=it)
        torch.cuda.synchronize()
        eW = float((tr.tensor("W") - ref.tensor("W")).abs().max() / ref.tensor("W").abs().max())
        eH = float((tr.tensor("W_hist") - ref.tensor("W_hist")).abs().max() / ref.tensor("W_hist").abs().max())
        eL = abs(tr.tensor("loss").item() - ref.tensor("loss").item())
        eV = abs(tr.tensor("violations").item() - ref.tensor("violations").item())
        worst = max(worst, eW, eH, eL)
        print("iter %d: relerr W %.2e hist %.2e loss %.2e viol %g" % (it, eW, eH, eL, eV))
# replicas must stay bit-identical across ranks
w = tr.tensor("W").clone(); wmax = w.clone(); wmin = w.clone()
dist.all_reduce(wmax, op=dist.ReduceOp.MAX); dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
same = bool(torch.equal(wmax, wmin))
if rank == 0:
    tol = 1e-5 if prec in ("tf32x3", "f16x3", "fp32_simt") else 5e-2
    print("DP_CHECK %s world=%d prec=%s fused_gather=%s worst=%.2e replicas_identical=%s" % ("OK" if (worst < tol and same) else "FAIL", world, prec, fused, worst, same))
dist.barrier()
dist.destroy_process_group()

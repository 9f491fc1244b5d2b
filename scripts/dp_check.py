"""Run under torchrun with G ranks: G-rank data-parallel steps vs a 1-rank trainer on the global batch.
Prints 'DP_CHECK OK ...' on rank 0 (used by tests/test_gpu_dp.py and the multi-GPU GPU trip).
  dp_check.py <precision> [steps]         VV_DP_MODE=nccl forces the NCCL all-reduce path (default: peer-memory exchange)
Checks: (1) W, history, bias, loss, violations of the G-rank run against one rank on the global batch (1e-5 for the
fp32-parity modes; the summation order differs), every step; (2) replicas bit-identical: the operand copy of W, W[:,K-1]
and the bias every rank ends up with; (3) HASH-mode dropout draws a different mask on every rank."""
import os, sys, threading, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from videovector_b200 import ops, dp
from videovector_b200._lib import DROPOUT_MASK01, DROPOUT_HASH

# a rank that fails must not leave the others waiting in a collective: any uncaught exception ends the process at once
# (torchrun then stops the other ranks), and a watchdog ends a run that outlives its budget
def _die(exc_type, exc, tb):
    traceback.print_exception(exc_type, exc, tb); sys.stdout.flush(); sys.stderr.flush(); os._exit(1)
sys.excepthook = _die
threading.Timer(float(os.environ.get("DP_CHECK_BUDGET_S", "120")), lambda: (print("DP_CHECK FAIL watchdog: run exceeded its budget", flush=True), os._exit(2))).start()

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B, C, Nn, K, N = 64, 5, 10, 1024, 512          # per rank
fused = prec in ("f16x3", "bf16") and os.environ.get("VV_DP_MATERIALISE") != "1"
R = C + Nn
V, S = 128, 24
vid, off, sid = ops.synthetic_videos(V, S)
bank = ops.fill_bank(V * S, K, 1234)
rng = np.random.RandomState(1701)
W0 = torch.as_tensor(rng.normal(0, 0.02, (N, K)).astype(np.float32)).cuda()
b0 = torch.as_tensor(rng.normal(0, 0.01, N).astype(np.float32)).cuda()
sol = dict(base_lr=0.05)


def make(world_size, rk, per_rank_B, mode=DROPOUT_MASK01):
    tr = ops.Trainer(ops.trainer_cfg(per_rank_B, C, Nn, K, N, prec=prec, dropout_ratio=0.5, dropout_mode=mode,
                                     world_size=world_size, rank=rk, **sol))
    tr.set_weights(W0, b0)
    if fused:
        tr.set_bank(bank)
    if world_size > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rk == 0:
            idt.copy_(torch.frombuffer(bytearray(ops.dp_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        tr.dp_init(bytes(idt.cpu().numpy().tobytes()))
    return tr


def same_everywhere(t):
    """bit-identical on every rank (compared as integers: NaN-proof)"""
    v = t.contiguous().view(torch.int32).clone()
    hi, lo = v.clone(), v.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(torch.equal(hi, lo))


tr = make(world, rank, B)
mode = tr.dp_mode
ref = make(1, 0, B * world) if rank == 0 else None
smp = ops.Sampler(vid, off, sid, B * world, C, Nn, 500, 50, 6, 100, rand_seed=1)     # global stream, same on every rank
mrng = np.random.RandomState(3)
worst = 0.0          # strict bound (1e-5 for the fp32-parity modes): every step that starts from identical state
drift = 0.0          # free-running steps: the runs may part at ReLU gates (see below)


def one_step(it, strict):
    """One update step on the G ranks and on the 1-rank reference (global batch), compared on rank 0."""
    global worst, drift
    gidx, gq = smp.next()
    gmask = (mrng.uniform(0, 1, (R, B * world, N)) > 0.5).astype(np.int32)
    idx, quirk = dp.shard_batch(gidx, gq, rank, world)
    mask = np.ascontiguousarray(gmask[:, rank * B:(rank + 1) * B]).reshape(R * B, N)
    tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), torch.as_tensor(mask).cuda(), it=it)
    tr.dp_gather_state()                      # collective: the owner-sharded master weights / history (p2p mode)
    if rank == 0:
        ref.step(bank, torch.as_tensor(gidx).cuda(), torch.as_tensor(gq).cuda(), torch.as_tensor(gmask.reshape(R * B * world, N)).cuda(), it=it)
        torch.cuda.synchronize()
        rel = lambda a: float((tr.tensor(a) - ref.tensor(a)).abs().max() / ref.tensor(a).abs().max())
        eW, eH, eb = rel("W"), rel("W_hist"), rel("b")
        eL = abs(tr.tensor("loss").item() - ref.tensor("loss").item())
        eV = abs(tr.tensor("violations").item() - ref.tensor("violations").item())
        if strict:
            worst = max(worst, eW, eH, eb, eL, 1.0 if eV != 0 else 0.0)
        else:
            drift = max(drift, eW, eL)
        print("iter %d (%s): relerr W %.2e hist %.2e b %.2e loss %.2e viol %g" % (it, "strict" if strict else "free", eW, eH, eb, eL, eV))


def resync():
    """Every rank (and nothing else changes) restarts from the reference's exact state: W, b and both histories."""
    bufs = [torch.empty((N, K), device="cuda"), torch.empty(N, device="cuda"), torch.empty((N, K), device="cuda"), torch.empty(N, device="cuda")]
    if rank == 0:
        for t_, name in zip(bufs, ("W", "b", "W_hist", "b_hist")):
            t_.copy_(ref.tensor(name))
    for t_ in bufs:
        dist.broadcast(t_, 0)
    tr.tensor("W_hist").copy_(bufs[2]); tr.tensor("b_hist").copy_(bufs[3])
    tr.set_weights(bufs[0], bufs[1])
    torch.cuda.synchronize()
    dist.barrier()


# (a) step 0 starts from identical state on both sides: strict.  (b) free-running steps: once the weights differ in the
# last bit, a pre-activation within rounding of zero may take the other side of its ReLU gate (with 15*B*G*N of them per
# step that happens every few steps; it changes a gradient row by its whole magnitude, ~1e-3 of max|dW|) -- the loss stays
# within 1e-5, the weights within 2e-3.  (c) strict again: every step restarts from the reference's exact state.
one_step(0, True)
for it in range(1, steps):
    one_step(it, False)
for it in range(steps, steps + 3):
    resync()
    one_step(it, True)
steps += 3
# one forward/backward-only step (the NCCL path in both modes): loss is the mean over ranks, gradients all-reduced
resync()
gidx, gq = smp.next()
gmask = (mrng.uniform(0, 1, (R, B * world, N)) > 0.5).astype(np.int32)
idx, quirk = dp.shard_batch(gidx, gq, rank, world)
mask = np.ascontiguousarray(gmask[:, rank * B:(rank + 1) * B]).reshape(R * B, N)
tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), torch.as_tensor(mask).cuda(), it=steps, do_update=False)
if rank == 0:
    ref.step(bank, torch.as_tensor(gidx).cuda(), torch.as_tensor(gq).cuda(), torch.as_tensor(gmask.reshape(R * B * world, N)).cuda(),
             it=steps, do_update=False)
    torch.cuda.synchronize()
    eL = abs(tr.tensor("loss").item() - ref.tensor("loss").item())
    eG = float((tr.tensor("dW_raw") / world - ref.tensor("dW_raw")).abs().max() / ref.tensor("dW_raw").abs().max())
    worst = max(worst, eL, eG)
    print("no-update step: loss %.2e dW %.2e" % (eL, eG))
# replicas must stay bit-identical across ranks: what every rank multiplies with in the next forward
torch.cuda.synchronize()
same = same_everywhere(tr.tensor("W")) and same_everywhere(tr.tensor("b")) and same_everywhere(tr.tensor("wlast"))
if prec in ("f16x3", "bf16", "tf32x3"):
    same = same and same_everywhere(tr.tensor("Wop_hi"))
    if prec != "bf16":
        same = same and same_everywhere(tr.tensor("Wop_lo"))
# HASH-mode dropout: identical inputs on every rank, the masks (hence H) must differ between ranks
th = make(world, rank, B, mode=DROPOUT_HASH)
i0, q0 = dp.shard_batch(gidx, gq, 0, world)
th.step(bank, torch.as_tensor(i0).cuda(), torch.as_tensor(q0).cuda(), None, it=0, do_update=False)
torch.cuda.synchronize()
masks_differ = not same_everywhere((th.tensor("H") != 0).to(torch.int32))
if rank == 0:
    tol = 1e-5 if prec in ("tf32x3", "f16x3", "fp32_simt") else 5e-2
    ok = worst < tol and drift < max(2e-3, tol) and same and masks_differ
    print("DP_CHECK %s world=%d prec=%s mode=%s%s fused_gather=%s worst=%.2e free_running_drift=%.2e replicas_identical=%s hash_masks_differ=%s" % (
        "OK" if ok else "FAIL", world, prec, mode, (" (" + tr.dp_mode_reason + ")") if tr.dp_mode_reason else "", fused, worst, drift, same,
        masks_differ))
dist.barrier()
th.close(); tr.close()
if ref is not None:
    ref.close()
dist.destroy_process_group()
os._exit(0)

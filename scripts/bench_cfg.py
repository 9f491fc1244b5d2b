"""Per-phase timing of one configuration (CUDA events inside the step) -- e.g. BASELINE configs[3]:
python scripts/bench_cfg.py --C 17 --Nn 50 --N 1024 --B 4096"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops
from videovector_b200._lib import DROPOUT_HASH

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="f16x3")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--C", type=int, default=5)
ap.add_argument("--Nn", type=int, default=10)
ap.add_argument("--N", type=int, default=512)
ap.add_argument("--K", type=int, default=4096)
ap.add_argument("--materialised", action="store_true")
ap.add_argument("--split-rank", action="store_true")
a = ap.parse_args()
torch.cuda.set_device(0)
B, C, Nn, K, N = a.B, a.C, a.Nn, a.K, a.N
V, S = 8192, 32
bank = ops.fill_bank(V * S, K, 1234)
vid, off, sid = ops.synthetic_videos(V, S)
smp = ops.Sampler(vid, off, sid, B, C, Nn, 5000, 50, 6, 100, rand_seed=1)
tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, prec=a.precision, dropout_mode=DROPOUT_HASH, split_rank_loss=a.split_rank))
tr.set_weights(torch.randn(N, K, device="cuda") * 0.001, torch.zeros(N, device="cuda"))
if a.precision in ("f16x3", "bf16") and not a.materialised:
    tr.set_bank(bank)
batches = [tuple(torch.as_tensor(x).cuda() for x in smp.next()) for _ in range(a.steps + 3)]
for it in range(3):
    tr.step(bank, batches[it][0], batches[it][1], None, it=it)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(a.steps):
    tr.step(bank, batches[3 + it][0], batches[3 + it][1], None, it=3 + it)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
tr.set_timing(True)
for it in range(min(a.steps, 10)):
    tr.step(bank, batches[3 + it][0], batches[3 + it][1], None, it=100 + it)
phase, _ = tr.phase_ms()
R = C + Nn; M = R * B
print("cfg C=%d Nn=%d N=%d B=%d %s: %.3f ms/step = %.0f triplets/s; loss %.4f" % (C, Nn, N, B, a.precision, ms, B / ms * 1e3, tr.tensor("loss").item()))
print("  phases (ms):", {k: round(v, 4) for k, v in phase.items() if v > 0})
fl = 2.0 * M * N * K
for k in ("fc7_forward", "wgrad"):
    if phase.get(k):
        print("  %s: %.0f TFLOP/s algorithmic" % (k, fl / phase[k] / 1e9))
hb = M * N * 4
for k, by in (("rank_loss_forward", hb), ("rank_loss_backward", hb * 2)):
    if phase.get(k):
        print("  %s: %.0f GB/s (fp32 read%s)" % (k, by / phase[k] / 1e6, " + operand write" if k.endswith("backward") else ""))

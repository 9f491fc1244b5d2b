"""A few fused training steps at BASELINE config[1] shapes for ncu (no timing, no CPU baseline)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops
from videovector_b200._lib import DROPOUT_HASH

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="f16x3")
ap.add_argument("--materialised", action="store_true")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--C", type=int, default=5)
ap.add_argument("--Nn", type=int, default=10)
ap.add_argument("--N", type=int, default=512)
a = ap.parse_args()
torch.cuda.set_device(0)
B, C, Nn, K, N = a.B, a.C, a.Nn, 4096, a.N
V, S = 8192, 32
bank = ops.fill_bank(V * S, K, 1234)
vid, off, sid = ops.synthetic_videos(V, S)
smp = ops.Sampler(vid, off, sid, B, C, Nn, 5000, 50, 6, 100, rand_seed=1)
tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, prec=a.precision, dropout_mode=DROPOUT_HASH))
tr.set_weights(torch.randn(N, K, device="cuda") * 0.001, torch.zeros(N, device="cuda"))
if a.precision in ("f16x3", "bf16") and not a.materialised:
    tr.set_bank(bank)
for it in range(a.steps):
    idx, quirk = smp.next()
    tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), None, it=it)
torch.cuda.synchronize()
print("loss", tr.tensor("loss").item(), "launches/step", tr.last_launches)

"""A few small training steps + the evaluation kernels for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from videovector_b200 import ops
from videovector_b200._lib import DROPOUT_HASH

torch.cuda.set_device(0)
prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
B, C, Nn, K, N = 24, 5, 10, 256, 512
V, S = 64, 24
bank = ops.fill_bank(V * S, K, 1234)
vid, off, sid = ops.synthetic_videos(V, S)
smp = ops.Sampler(vid, off, sid, B, C, Nn, 200, 50, 6, 100, rand_seed=1)
tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, prec=prec, dropout_mode=DROPOUT_HASH, compute_dgrad=True))
tr.set_weights(torch.randn(N, K, device="cuda") * 0.02, torch.zeros(N, device="cuda"))
if prec in ("f16x3", "bf16") and len(sys.argv) <= 2:
    tr.set_bank(bank)
for it in range(3):
    idx, quirk = smp.next()
    tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), None, it=it)
torch.cuda.synchronize()
print("steps ok, loss", tr.tensor("loss").item())
E = torch.nn.functional.normalize(torch.randn(100, 64, device="cuda"))
out = ops.retrieval_stats(E, np.arange(100) % 17, np.arange(100) % 5, True)
ids = torch.as_tensor((np.arange(50) % 7).astype(np.float32)).cuda()
d = ops.id_lookup_backward(torch.randn(50, 32, device="cuda"), ids, 9)
torch.cuda.synchronize()
print("eval ok", out["map"])

"""Times the K2+K3 kernel for wide items (configs[3] shape: 17 context rows incl. target + 50 negatives, N = 1024) alone:
H is 1.1 GB (>> L2), operand-only f16x3 output.  Variants by environment (VV_RANK_WIDE, VV_RANK_WIDE_CTAS,
VV_RANK_WIDE_PREFETCH) -- one process per variant, the launcher reads them once."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--C", type=int, default=17)
ap.add_argument("--Nn", type=int, default=50)
ap.add_argument("--N", type=int, default=1024)
ap.add_argument("--prec", default="f16x3")
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
torch.cuda.set_device(0)
R = a.C + a.Nn
g = torch.Generator(device="cuda").manual_seed(1)
H = torch.relu(torch.randn(R * a.B, a.N, device="cuda", generator=g)).contiguous()
cfg = ops.rank_cfg(a.B, a.C, a.Nn, a.N, margin=2.0, norm=2)
fused = ops.rank_loss_fused_supported(cfg)
assert fused, "VV_RANK_WIDE=0: the two-kernel path is timed by bench.py (rank_loss_forward + rank_loss_backward)"
import ctypes as C
from videovector_b200 import _lib
from videovector_b200.ops import _ptr, _stream, _prec, alloc_operand, operand_rescale, check
p = _prec(a.prec)
dev = H.device
stats = torch.empty((a.B, 1 + 2 * (1 + a.Nn)), dtype=torch.float32, device=dev)
item_loss = torch.empty((a.B,), dtype=torch.float32, device=dev); item_viol = torch.empty_like(item_loss)
loss = torch.empty((1,), dtype=torch.float32, device=dev); viol = torch.empty_like(loss)
db = torch.zeros((a.N,), dtype=torch.float32, device=dev)
op = alloc_operand(H.shape, p, dev)
def once():
    check(_lib.load().vv_rank_loss_fused(_ptr(H), C.byref(cfg), 1.0, 1, 10.0, _ptr(stats), None, None, _ptr(item_loss), _ptr(item_viol),
                                         _ptr(loss), _ptr(viol), None, _ptr(op.hi), _ptr(op.lo) if op.lo is not None else None, p,
                                         _ptr(db), None, None, _stream()))
once()
if a.prec == "f16x3":
    operand_rescale(op, p)
for _ in range(3): once()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters): once()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters                     # includes the 2 us batch-loss reduce launch (no workspace through the C-ABI)
opb = {"f16x3": 4, "bf16": 2, "tf32x3": 8}[a.prec]
gb = R * a.B * a.N * (4 + opb) / 1e9
print("env", {k: v for k, v in os.environ.items() if k.startswith("VV_RANK")}, "ms %.4f" % ms, "alg GB %.3f" % gb,
      "GB/s %.0f" % (gb / ms * 1e3), "loss", float(loss))

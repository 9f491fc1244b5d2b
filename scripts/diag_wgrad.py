"""Bring-up probe for the weight-gradient GEMMs (materialised operands): relative error and, when it is large, the error per
128 x 256 output tile, for every tensor-core precision at one tile, a few tiles and cfg-1's shape, with 1 and 2 split-K slabs.
(Found the cluster-rank bit leaking into the MN-major descriptors' LBO field, DESIGN.md section 3, "The MMA issuer".)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops
torch.manual_seed(0)
for (M, N, K) in [(128, 256, 64), (256, 256, 256), (1920, 512, 4096)]:
    X = torch.randn(M, K, device="cuda"); dZ = torch.randn(M, N, device="cuda")
    ref = dZ.double().t() @ X.double()
    for prec in ("bf16", "tf32", "f16x3", "tf32x3"):
        for ns in (1, 2):
            if ns > 1 and M < 256: continue
            try:
                dW = ops.ip_wgrad(ops.prepare_operand(dZ, prec), ops.prepare_operand(X, prec), M, N, K, prec, nsplit=ns)
                torch.cuda.synchronize()
            except Exception as e:
                print(M, N, K, prec, ns, "ERROR", str(e)[:200]); continue
            err = (dW.double() - ref).abs()
            r = float(err.max() / ref.abs().max())
            # error per 128 x 256 tile of dW [N, K]
            tiles = err.reshape(max(N // 128, 1), min(128, N), max(K // 256, 1), min(256, K)).amax(dim=(1, 3)) if (N % 128 == 0 and K % 256 == 0) else None
            print(M, N, K, prec, "nsplit", ns, "rel %.2e" % r, "ratio of mean |dW|/|ref| %.3f" % float(dW.abs().mean() / ref.abs().mean()),
                  "" if tiles is None or r < 1e-2 else (tiles / ref.abs().max()).cpu().numpy().round(2).tolist()[:2])

"""BASELINE configs[4] slice: embedding inference relu(F W^T + b) on one GPU's shard of rows (tools/extract_features.cpp:
100-209, blob ip2), timed per precision.  Algorithmic work per row: 2*K*N flops, K*4 + N*4 bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops

torch.cuda.set_device(0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K, N = 4096, 512
F = ops.fill_bank(rows, K, 77)
W = torch.randn(N, K, device="cuda") * 0.01; b = torch.zeros(N, device="cuda")
for prec in ("tf32", "bf16", "f16x3"):
    tr = ops.Trainer(ops.trainer_cfg(4096, 5, 10, K, N, prec=prec))
    tr.set_weights(W, b)
    out = tr.extract(F); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = tr.extract(F); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-6s %d rows: %.2f ms = %.2f M rows/s; %.0f TFLOP/s algorithmic; %.0f GB/s algorithmic (%.0f%% of 6552)" % (
        prec, rows, ms, rows / ms / 1e3, 2.0 * rows * K * N / ms / 1e9, rows * (K + N) * 4 / ms / 1e6, rows * (K + N) * 4 / ms / 1e6 / 65.52))
    tr.close()

"""Blob-by-blob comparison of the reference pipeline in Caffe CPU mode (libvv_ref.so) and in GPU mode over the C-ABI
(libvv_dropin.so) after one solver step on the trajectory fixture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyref
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "solver_ref.npz"))
B, C, Nn, P, swap, max_same = [int(x) for x in g["cfg"]]
base_lr, mom, wd, gamma, power = [float(x) for x in g["hyper"]]
def mk(lib):
    return pyref.Solver(g["vid"], g["off"], g["sid"], g["feat"], g["W0"], g["b0"], B, C, Nn, P, swap, max_same, base_lr=base_lr,
                        momentum=mom, weight_decay=wd, lr_policy="inv", gamma=gamma, power=power, library=lib)
def run(lib):
    sol = mk(lib)
    res = sol.step()
    blobs = {(n, d): sol.blob(n, d) for n in sol.blob_names() for d in (False, True)}
    st = sol.state(); names = sol.blob_names(); sol.close()
    return res, blobs, st, names
ra, A, sa, names = run(None)                      # one solver at a time: the CPU build draws from libc's global rand()
rb, Bb, sb, _ = run(pyref.dropin_lib())
print("cpu step", ra, "gpu step", rb)
shown = 0
for name in names:
    for diff in (False, True):
        x, y = A[(name, diff)], Bb[(name, diff)]
        e = float(np.abs(x - y).max() / max(np.abs(x).max(), 1e-30)) if x.size == y.size else -1
        if (e > 1e-5 or e < 0):
            shown += 1
            print("%s.%s:%.2e(max %.1e, gpu max %.1e)" % (name, "diff" if diff else "data", e, np.abs(x).max(), np.abs(y).max()), end="  ")
print()
for k in ("W", "b", "hW", "hb"):
    print(k, float(np.abs(sa[k] - sb[k]).max() / np.abs(sa[k]).max()))

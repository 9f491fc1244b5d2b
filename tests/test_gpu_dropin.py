"""The drop-in proof: the reference's OWN Net / SGDSolver / Blob / SyncedMemory and data layer -- compiled from its sources
in Caffe GPU mode against its real headers (oracle/ref_shim/build_ref.sh -> oracle/_ref/libvv_dropin.so) -- drive the new
device code: every symbol the reference's .cu files would define (the layers' Forward_gpu / Backward_gpu of
common_layers.hpp / neuron_layers.hpp / data_layers.hpp, util/math_functions.cu's caffe_gpu_*) is a call sequence into
libvv_b200.so's C-ABI (oracle/ref_shim/dropin_gpu.cpp).  The trajectories must be the ones the same sources produced in
Caffe CPU mode (tests/golden/solver_ref.npz): loss and violations per iteration, final weights, bias and both histories.
(ref: include/caffe/layer.hpp:25-404, net.cpp:504-581, solver.cpp:177-220,486-576.)  Skipped where the library was not
built (it needs /root/reference at build time; the built file travels with the tree)."""
import os

import numpy as np
import pytest

from oracle import pyref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("prec,tol", [("f16x3", 1e-5), ("tf32x3", 1e-5), ("fp32_simt", 1e-5)])
def test_reference_net_and_solver_drive_the_new_device_code(prec, tol):
    if not pyref.dropin_available():
        pytest.skip("oracle/_ref/libvv_dropin.so was not built (needs /root/reference at build time)")
    import subprocess, sys, json
    # one process per precision: the precision is read once (VV_DROPIN_PREC) and the Caffe singleton is process-wide
    code = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, %r)
from oracle import pyref
g = np.load(%r)
B, C, Nn, P, swap, max_same = [int(x) for x in g["cfg"]]
base_lr, mom, wd, gamma, power = [float(x) for x in g["hyper"]]
L = pyref.dropin_lib()
sol = pyref.Solver(g["vid"], g["off"], g["sid"], g["feat"], g["W0"], g["b0"], B, C, Nn, P, swap, max_same, base_lr=base_lr,
                   momentum=mom, weight_decay=wd, lr_policy="inv", gamma=gamma, power=power, library=L)
names = sol.layer_names()
traj = [sol.step() for _ in range(len(g["loss"]))]
st = sol.state(); sol.close()
a_lr, a_mom, a_wd, a_gamma, a_step = [float(x) for x in g["alt_hyper"]]
sol = pyref.Solver(g["vid"], g["off"], g["sid"], g["feat"], g["W0"], g["b0"], B, C, Nn, P, swap, max_same, norm=1, reg_type=1,
                   base_lr=a_lr, momentum=a_mom, weight_decay=a_wd, lr_policy="step", gamma=a_gamma, power=0.0, stepsize=int(a_step),
                   library=L)
alt = [sol.step() for _ in range(len(g["alt_loss"]))]
sa = sol.state(); sol.close()
np.savez(sys.argv[1], loss=[t[0] for t in traj], viol=[t[1] for t in traj], W=st["W"], b=st["b"], hW=st["hW"], hb=st["hb"],
         alt_loss=[t[0] for t in alt], alt_viol=[t[1] for t in alt], alt_W=sa["W"], alt_b=sa["b"], alt_hW=sa["hW"], alt_hb=sa["hb"],
         names=np.array(names))
''' % (ROOT, os.path.join(GOLD, "solver_ref.npz"))
    out = os.path.join(os.environ.get("TMPDIR", "/tmp"), "vv_dropin_%s_%d.npz" % (prec, os.getpid()))
    r = subprocess.run([sys.executable, "-c", code, out], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, VV_DROPIN_PREC=prec))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    got = np.load(out); os.remove(out)
    g = np.load(os.path.join(GOLD, "solver_ref.npz"))
    assert list(got["names"]) == list(g["layer_names"])              # the reference's own Net::Init built the graph
    for it in range(len(g["loss"])):
        assert abs(got["loss"][it] - g["loss"][it]) < tol * max(1, abs(g["loss"][it])), (it, got["loss"][it], g["loss"][it])
        assert got["viol"][it] == g["violations"][it], it
    for k in ("W", "b", "hW", "hb"):
        assert rel(got[k], g[k]) < 2 * tol, (k, rel(got[k], g[k]))
    for it in range(len(g["alt_loss"])):                             # L1 hinge, step policy, L1 regularisation
        assert abs(got["alt_loss"][it] - g["alt_loss"][it]) < tol * max(1, abs(g["alt_loss"][it])), it
        assert got["alt_viol"][it] == g["alt_violations"][it], it
    for k in ("W", "b", "hW", "hb"):
        assert rel(got["alt_" + k], g["alt_" + k]) < 2 * tol, k

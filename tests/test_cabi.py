"""The C-ABI library loads on a CPU-only box and exports every symbol include/vv_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "vv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(vv_[a-z0-9_]+)\s*\(", src)
    # drop the static inline helper defined in the header itself
    return sorted(set(n for n in names if n != "vv_rank_stats_stride"))


def test_header_declares_functions():
    names = header_functions()
    assert len(names) > 40
    for must in ("vv_ip_forward", "vv_ip_wgrad", "vv_ip_dgrad", "vv_rank_loss_forward", "vv_rank_loss_backward",
                 "vv_sgd_update", "vv_gather_rows", "vv_sampler_next", "vv_trainer_step"):
        assert must in names


def test_library_exports_every_declared_symbol(vvlib):
    missing = [n for n in header_functions() if not hasattr(vvlib, n)]
    assert not missing, "declared in vv_b200.h but not exported: %s" % missing


def test_python_binding_covers_header(vvlib):
    from videovector_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()


def test_no_cpu_fallback_without_gpu(vvlib):
    """Host-only entry points work; anything touching the device fails loudly (never falls back)."""
    import torch
    assert vvlib.vv_version() >= 100
    if not torch.cuda.is_available():
        assert vvlib.vv_device_check() != 0
        assert len(vvlib.vv_last_error()) > 0


def test_product_does_not_reference_oracle():
    """The product tree must not import, link or call anything under oracle/."""
    bad = []
    for base in ("videovector_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"pyoracle|vv_oracle|orc_[a-z_]+\(|from oracle|import oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_learning_rate_host(vvlib, oracle):
    for it in (0, 1, 10, 999, 200000):
        a = vvlib.vv_learning_rate(b"inv", 1e-3, 1e-3, 0.75, 1, it)
        b = oracle.learning_rate("inv", 1e-3, 1e-3, 0.75, 1, it)
        assert a == b
    assert vvlib.vv_learning_rate(b"step", 0.1, 0.5, 0.0, 10, 35) == oracle.learning_rate("step", 0.1, 0.5, 0.0, 10, 35)
    assert vvlib.vv_learning_rate(b"exp", 0.1, 0.99, 0.0, 1, 7) == oracle.learning_rate("exp", 0.1, 0.99, 0.0, 1, 7)
    assert vvlib.vv_learning_rate(b"fixed", 0.1, 0.99, 0.0, 1, 7) == oracle.learning_rate("fixed", 0.1, 0.99, 0.0, 1, 7)

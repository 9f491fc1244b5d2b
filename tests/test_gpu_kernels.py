"""Parity of the sm_100a kernels (through the C-ABI) against the CPU oracle and float64 references.

Tolerances (north_star): indices / gathers bit-exact; fp32 paths (FP32_SIMT, TF32X3) within 1e-5
relative of the oracle; TF32 / BF16 tensor-core paths within the stated loose bounds below.
Relative error is max|a-b| / max|b| over the tensor (BLAS summation order is unspecified)."""
import numpy as np
import pytest
import torch

from videovector_b200 import ops
from videovector_b200._lib import DROPOUT_HASH, DROPOUT_MASK01, DROPOUT_MASK_U32, DROPOUT_NONE, DROPOUT_PHILOX

pytestmark = pytest.mark.gpu

TOL = {"fp32_simt": 1e-5, "tf32x3": 1e-5, "f16x3": 1e-5, "tf32": 4e-3, "bf16": 2e-2}
TC = ["tf32x3", "f16x3", "tf32", "bf16"]
ALL = ["fp32_simt"] + TC


def rel(a, b):
    a = a.double() if torch.is_tensor(a) else torch.as_tensor(a).double()
    b = b.double() if torch.is_tensor(b) else torch.as_tensor(b).double()
    a, b = a.cpu(), b.cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_x3_operand(op, x):
    """TF32X3 operand format: hi = tf32(x) (13 low mantissa bits clear); lo = two bf16 planes [bf16(x) | bf16(x - hi)]."""
    n = x.numel()
    assert torch.equal(op.hi.view(torch.int32) & 0x1FFF, torch.zeros_like(op.hi, dtype=torch.int32))
    assert rel(op.hi, x) < 2 ** -11
    planes = op.lo.reshape(-1).view(torch.bfloat16)
    assert planes.numel() == 2 * n
    assert torch.equal(planes[:n], x.reshape(-1).to(torch.bfloat16))
    assert torch.equal(planes[n:], (x - op.hi).reshape(-1).to(torch.bfloat16))
    assert rel(op.hi.double().reshape(-1) + planes[n:].double(), x.reshape(-1)) < 2 ** -19


def check_f16x3_operand(op, x):
    """F16X3 operand: fp16 planes h0 = fp16(s*x), h1 = fp16(s*x - h0) under the header's power-of-two scale s; the
    producer records max|x|; h0 + h1 carries ~22 significant bits (less only for elements far below the maximum)."""
    s = op.scale
    assert s > 0 and np.log2(s) == int(np.log2(s))
    assert op.absmax == float(x.abs().max().item())
    xs = x.double() * s
    assert float(xs.abs().max()) < 65504                                    # nothing saturates
    assert torch.equal(op.hi, (x * s).to(torch.float16))
    assert torch.equal(op.lo, (x * s - op.hi.float()).to(torch.float16))
    err = (op.hi.double() + op.lo.double() - xs).abs()
    # 22 significant bits, down to an absolute floor where h1 turns subnormal (2^-25 in scaled units)
    assert bool((err <= torch.clamp(xs.abs() * 2.0 ** -21.9, min=2.0 ** -25)).all())


def cuda(a, dtype=None):
    t = torch.as_tensor(a)
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


@pytest.fixture(scope="module", autouse=True)
def _device(vvlib):
    assert vvlib.vv_device_check() == 0, vvlib.vv_last_error()
    torch.cuda.set_device(0)


# ------------------------------------------------------------------------------------------------
# K0 gather + synthetic bank: bit-exact
# ------------------------------------------------------------------------------------------------
def test_bank_fill_bit_exact():
    bank = ops.fill_bank(257, 64, 1234)
    assert np.array_equal(bank.cpu().numpy(), ops.bank_host(257, 64, 1234))


@pytest.mark.parametrize("prec", ALL)
def test_gather_rows_bit_exact(oracle, prec):
    rng = np.random.RandomState(5)
    K, B, C, Nn = 64, 16, 5, 10
    from test_sampler import make_dataset
    video_id, shot_off, shot_ids, feat = make_dataset(rng, 80, 4, 20, K)
    osmp = oracle.Sampler(video_id, shot_off, shot_ids, feat, K, B, C, Nn, 60, 50, 6, 100, seed=1)
    batches = [osmp.next() for _ in range(4)]
    osmp.close()
    psmp = ops.Sampler(video_id, shot_off, shot_ids, B, C, Nn, 60, 50, 6, 100, rand_seed=1)
    bank = cuda(feat)
    for (oi, oq, data) in batches:
        idx, quirk = psmp.next()
        assert np.array_equal(idx, oi) and np.array_equal(quirk, oq)          # index stream bit-exact
        X, op, blob = ops.gather_rows(bank, cuda(idx), cuda(quirk), prec, want_x=True, want_blob=True)
        assert np.array_equal(blob.cpu().numpy(), data)                       # the data blob, K-1 quirk included
        Xref = np.ascontiguousarray(data.transpose(1, 0, 2)).reshape((C + Nn) * B, K)   # slice dim1 + concat dim0
        assert np.array_equal(X.cpu().numpy(), Xref)
        if prec == "bf16":
            assert torch.equal(op.hi, X.to(torch.bfloat16))
        if prec == "tf32x3":
            check_x3_operand(op, X)
        if prec == "f16x3":
            assert op.absmax == float(X.abs().max().item())     # recorded by the gather; the scale came from max|bank|
            s = op.scale
            assert float(bank.abs().max()) * s < 65504 and torch.equal(op.hi, (X * s).to(torch.float16))
            assert torch.equal(op.lo, (X * s - op.hi.float()).to(torch.float16))
    psmp.close()


def test_gather_no_quirk_pointer():
    bank = ops.fill_bank(100, 32, 9)
    idx = cuda(np.random.RandomState(0).randint(0, 100, (8, 6)).astype(np.int32))
    X, _, _ = ops.gather_rows(bank, idx, None)
    ref = bank[idx.long().t().reshape(-1)]
    assert torch.equal(X, ref)


# ------------------------------------------------------------------------------------------------
# K1 GEMMs
# ------------------------------------------------------------------------------------------------
SHAPES = [(1920, 512, 4096),   # cfg-1: B=128, R=15
          (128, 256, 64),      # exactly one tile, two k-blocks (bf16) / two (tf32)
          (130, 264, 72),      # ragged M, N, K tails (TMA zero fill + predicated stores)
          (200, 24, 40),       # smaller than one tile in every dimension
          (1005, 1024, 512),   # cfg-4 embedding width
          # > 4 work units per CTA of the persistent kernels (157 m-tile pairs x 2 n-tiles on 74 clusters) with an ODD number
          # of 512-element promotion chunks per unit (K = 1280 = 2.5 chunks): the unit-to-unit hand-off -- the epilogue of tile
          # i under the main loop of tile i+1, the stage ring and both TMEM buffers carried across units -- against fp64
          (40000, 512, 1280)]


def _inputs(M, N, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    X = torch.relu(torch.randn(M, K, device="cuda", generator=g))
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g) * 0.1
    dZ = torch.randn(M, N, device="cuda", generator=g) * 0.01
    return X, W, b, dZ


@pytest.mark.parametrize("prec", ALL)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_ip_forward_plain(prec, M, N, K):
    X, W, b, _ = _inputs(M, N, K)
    H, _ = ops.ip_forward(ops.prepare_operand(X, prec), ops.prepare_operand(W, prec), b, M, N, K, prec)
    ref = X.double() @ W.double().t() + b.double()
    assert rel(H, ref) < TOL[prec], (prec, rel(H, ref))


@pytest.mark.parametrize("prec", ALL)
def test_ip_forward_matches_oracle_cfg1(oracle, prec):
    M, N, K = 1920, 512, 4096
    X, W, b, _ = _inputs(M, N, K, seed=3)
    oracle.use_openblas(0)
    ref = oracle.ip_forward(X.cpu().numpy(), W.cpu().numpy(), b.cpu().numpy())
    oracle.use_builtin_blas()
    H, _ = ops.ip_forward(ops.prepare_operand(X, prec), ops.prepare_operand(W, prec), b, M, N, K, prec)
    assert rel(H, ref) < TOL[prec]


@pytest.mark.parametrize("prec", ALL)
@pytest.mark.parametrize("mode", [DROPOUT_NONE, DROPOUT_MASK01, DROPOUT_MASK_U32, DROPOUT_PHILOX, DROPOUT_HASH])
def test_ip_forward_fused_relu_dropout(oracle, prec, mode):
    M, N, K = 256, 264, 96
    ratio = 0.9
    X, W, b, _ = _inputs(M, N, K, seed=1)
    g = torch.Generator(device="cuda").manual_seed(11)
    raw = torch.randint(0, 2 ** 32, (M, N), device="cuda", generator=g, dtype=torch.int64)   # u32 draws like curandGenerate
    thres = oracle.lib().orc_dropout_uint_thres(np.float32(ratio))
    mask = mask_out = None
    keep = (raw > thres).to(torch.int32).contiguous()
    if mode == DROPOUT_MASK01:
        mask = keep
    elif mode == DROPOUT_MASK_U32:
        mask = torch.where(raw >= 2 ** 31, raw - 2 ** 32, raw).to(torch.int32).contiguous()   # raw u32 bit patterns
    elif mode in (DROPOUT_PHILOX, DROPOUT_HASH):
        mask_out = torch.zeros((M, N), dtype=torch.int32, device="cuda")
        keep = ops.dropout_make_mask(M, N, ratio, 7, 5, mode=mode)
    act = ops.make_act(relu=True, dropout_mode=mode, ratio=ratio, mask=mask, mask_out=mask_out, seed=7, step=5)
    H, Z = ops.ip_forward(ops.prepare_operand(X, prec), ops.prepare_operand(W, prec), b, M, N, K, prec, act=act, want_z=True)
    Zref = oracle.ip_forward(X.cpu().numpy(), W.cpu().numpy(), b.cpu().numpy())
    assert rel(Z, Zref) < TOL[prec]
    # activation applied to the kernel's own Z must follow the reference layers exactly
    Href = oracle.relu_forward(Z.cpu().numpy())
    if mode != DROPOUT_NONE:
        Href = oracle.dropout_forward(Href, keep.cpu().numpy().astype(np.uint32), ratio)
        if mode in (DROPOUT_PHILOX, DROPOUT_HASH):
            assert torch.equal(mask_out, keep)
            frac = keep.float().mean().item()
            assert abs(frac - (1 - ratio)) < 0.01
    assert np.array_equal(H.cpu().numpy(), Href)


@pytest.mark.parametrize("mode", [DROPOUT_PHILOX, DROPOUT_HASH])
def test_generated_dropout_streams_are_sound(mode):
    """Both generated streams: keep rate = 1 - ratio to 3 sigma, rows / columns / steps / seeds decorrelated, and
    reproducible (same key -> same mask)."""
    rows, cols, ratio = 2048, 512, 0.9
    m = ops.dropout_make_mask(rows, cols, ratio, 7, 5, mode=mode).float()
    n = rows * cols
    assert abs(m.mean().item() - 0.1) < 3 * (0.09 / n) ** 0.5
    assert torch.equal(m, ops.dropout_make_mask(rows, cols, ratio, 7, 5, mode=mode).float())
    for other in (ops.dropout_make_mask(rows, cols, ratio, 7, 6, mode=mode), ops.dropout_make_mask(rows, cols, ratio, 8, 5, mode=mode)):
        o = other.float()
        both = (m * o).mean().item()                         # independent masks overlap on ~1 % of the entries
        assert abs(both - 0.01) < 5 * (0.01 / n) ** 0.5
    # neighbouring rows and columns are uncorrelated; per-row and per-column keep counts spread like a binomial
    assert abs(((m[1:] * m[:-1]).mean() - 0.01).item()) < 5 * (0.01 / n) ** 0.5
    assert abs(((m[:, 1:] * m[:, :-1]).mean() - 0.01).item()) < 5 * (0.01 / n) ** 0.5
    assert abs(m.sum(1).var().item() / (cols * 0.09) - 1) < 0.15 and abs(m.sum(0).var().item() / (rows * 0.09) - 1) < 0.25


@pytest.mark.parametrize("prec", ALL)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_ip_wgrad(prec, M, N, K):
    X, W, b, dZ = _inputs(M, N, K, seed=2)
    dW = ops.ip_wgrad(ops.prepare_operand(dZ, prec), ops.prepare_operand(X, prec), M, N, K, prec)
    ref = dZ.double().t() @ X.double()
    assert rel(dW, ref) < TOL[prec], (prec, rel(dW, ref))


@pytest.mark.parametrize("prec", ALL)
@pytest.mark.parametrize("nsplit", [1, 2, 3])
def test_ip_wgrad_explicit_splits_and_regularization(oracle, prec, nsplit):
    M, N, K = 1920, 512, 4096
    X, W, b, dZ = _inputs(M, N, K, seed=4)
    dW = ops.ip_wgrad(ops.prepare_operand(dZ, prec), ops.prepare_operand(X, prec), M, N, K, prec,
                      regularization=0.5, nsplit=nsplit)
    oracle.use_openblas(0)
    ref, db_ref, _ = oracle.ip_backward(dZ.cpu().numpy(), X.cpu().numpy(), W.cpu().numpy(), regularization=0.5)
    oracle.use_builtin_blas()
    assert rel(dW, ref) < TOL[prec]
    db = ops.ip_bias_grad(dZ)
    assert rel(db, db_ref) < 1e-5


@pytest.mark.parametrize("prec", ALL)
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_ip_dgrad(prec, M, N, K):
    X, W, b, dZ = _inputs(M, N, K, seed=6)
    dX = ops.ip_dgrad(ops.prepare_operand(dZ, prec), ops.prepare_operand(W, prec), M, N, K, prec)
    ref = dZ.double() @ W.double()
    assert rel(dX, ref) < TOL[prec], (prec, rel(dX, ref))


def test_gemm_known_answers_on_device():
    """The reference's GemmTest integers (test_util_blas.cpp:26-29) through the exact fp32 kernel."""
    A = cuda(np.array([[1, 2, 3], [4, 5, 6]], np.float32))
    Bm = cuda(np.arange(1, 13, dtype=np.float32).reshape(3, 4))
    # C = A * B  as  X W^T with W = B^T
    H, _ = ops.ip_forward(ops.OperandT(A), ops.OperandT(Bm.t().contiguous()), None, 2, 4, 3, "fp32_simt")
    assert H.flatten().tolist() == [38, 44, 50, 56, 83, 98, 113, 128]


# ------------------------------------------------------------------------------------------------
# K2 / K3 fused rank loss vs the oracle net and vs float64 autograd
# ------------------------------------------------------------------------------------------------
def _rank_ref64(H, B, C, Nn, margin, norm, lw, dscale):
    """float64 autograd on the GPU: d loss / d H and the fused-activation dZ."""
    R = C + Nn
    Hd = H.double().clone().requires_grad_(True)
    Hs = Hd.reshape(R, B, -1)
    a = float(np.float32(1.0 / (C - 1)))
    cbar = sum(Hs[i] * a for i in range(1, C))

    def l2n(x):
        s = x.pow(2).sum(1, keepdim=True)
        s = torch.where(s > 0, s, torch.ones_like(s))
        return x / (s.sqrt() + 1e-10)
    chat = l2n(cbar)
    st = (chat * l2n(Hs[0])).sum(1, keepdim=True)
    sn = torch.stack([(chat * l2n(Hs[C + k])).sum(1) for k in range(Nn)], 1)
    h = torch.clamp(margin - (st - sn), min=0)
    loss = (h.pow(2) if norm == 2 else h.abs()).sum() / (B * Nn)
    (loss * lw).backward()
    dZ = Hd.grad * dscale * (H > 0)
    return loss.item(), float((st - sn < 0).sum().item()), st.detach(), sn.detach(), Hd.grad, dZ


RANK_CASES = [(128, 5, 10, 512, 2), (64, 17, 50, 1024, 2), (37, 3, 4, 64, 1), (16, 5, 10, 4096, 2), (9, 7, 5, 200, 2)]


@pytest.mark.parametrize("B,C,Nn,N,norm", RANK_CASES)
def test_rank_loss_forward_backward(B, C, Nn, N, norm):
    R = C + Nn
    g = torch.Generator(device="cuda").manual_seed(B)
    H = torch.relu(torch.randn(R * B, N, device="cuda", generator=g))
    H = H * (torch.rand(R * B, N, device="cuda", generator=g) < 0.3) * 3.0      # sparse like dropout
    H[5] = 0.0                                                                   # an all-zero row
    H = H.contiguous()
    cfg = ops.rank_cfg(B, C, Nn, N, margin=2.0, norm=norm)
    out = ops.rank_loss_forward(H, cfg)
    loss, viol, st, sn, dH_ref, dZ_ref = _rank_ref64(H, B, C, Nn, 2.0, norm, 1.0, 10.0)
    assert abs(out["loss"].item() - loss) < 1e-5 * max(1, abs(loss))
    assert out["violations"].item() == viol
    assert rel(out["target_score"], st.expand(-1, Nn)) < 1e-5 and rel(out["neg_score"], sn) < 1e-5
    dH, _, _ = ops.rank_loss_backward(H, cfg, out["stats"], 1.0, act_fused=False, want_db=False)
    # the all-zero row: the reference's Normalization backward gives 0 there (s*dy - x*a = 0), the true
    # derivative of x/(|x|+eps) does not; with the ReLU/dropout gate fused (below) both are 0.
    assert dH[5].abs().max().item() == 0.0
    dH_ref[5] = 0.0
    assert rel(dH, dH_ref) < 1e-5, rel(dH, dH_ref)
    dZ, _, db = ops.rank_loss_backward(H, cfg, out["stats"], 1.0, act_fused=True, dropout_scale=10.0)
    assert rel(dZ, dZ_ref) < 1e-5
    assert rel(db, dZ_ref.sum(0)) < 1e-5


@pytest.mark.parametrize("B,C,Nn,N,norm", [(64, 5, 10, 512, 2), (33, 3, 4, 64, 1), (7, 5, 10, 1000, 2), (300, 7, 20, 256, 2),
                                            (1, 5, 10, 512, 2)])
@pytest.mark.parametrize("prec", ["fp32_simt", "tf32x3", "f16x3", "bf16"])
@pytest.mark.parametrize("ring", ["0", "2", "3"])
def test_rank_loss_fused_equals_two_kernel_path(B, C, Nn, N, norm, prec, ring, monkeypatch):
    """K2+K3 fused (rows in registers; ring != 0: the variant that stages items in shared memory with bulk async copies,
    VV_RANK_RING) against the forward + backward kernels: same formulas, reductions equal up to FMA contraction /
    summation order (1e-6)."""
    monkeypatch.setenv("VV_RANK_RING", ring)
    R = C + Nn
    g = torch.Generator(device="cuda").manual_seed(B + N)
    H = torch.relu(torch.randn(R * B, N, device="cuda", generator=g))
    H = (H * (torch.rand(R * B, N, device="cuda", generator=g) < 0.3) * 3.0).contiguous()
    H[min(5, R * B - 1)] = 0.0
    cfg = ops.rank_cfg(B, C, Nn, N, margin=2.0, norm=norm)
    assert ops.rank_loss_fused_supported(cfg)
    a = ops.rank_loss_forward(H, cfg)
    dZ_a, op_a, db_a = ops.rank_loss_backward(H, cfg, a["stats"], 0.7, True, 10.0, prec=prec)
    b, dZ_b, op_b, db_b = ops.rank_loss_fused(H, cfg, 0.7, True, 10.0, prec=prec)
    for k in ("stats", "target_score", "neg_score"):
        assert rel(b[k], a[k]) < 1e-6, k
    for k in ("item_viol", "violations"):
        assert torch.equal(a[k], b[k]), k
    assert rel(b["item_loss"], a["item_loss"]) < 1e-6 and rel(b["loss"], a["loss"]) < 1e-6
    assert rel(dZ_b, dZ_a) < 1e-6
    assert rel(db_b, db_a) < 1e-5                      # atomics: order differs
    if torch.equal(dZ_a, dZ_b) and prec != "fp32_simt":
        assert torch.equal(op_a.hi, op_b.hi) and (op_a.lo is None or torch.equal(op_a.lo, op_b.lo))
    loss, viol, st, sn, dH_ref, dZ_ref = _rank_ref64(H, B, C, Nn, 2.0, norm, 0.7, 10.0)
    assert abs(b["loss"].item() - loss) < 1e-5 * max(1, abs(loss)) and b["violations"].item() == viol
    assert rel(dZ_b, dZ_ref) < 1e-5


@pytest.mark.parametrize("B,C,Nn,N,norm", [(64, 5, 10, 512, 2), (33, 3, 4, 64, 1), (7, 5, 10, 1000, 2), (300, 7, 9, 256, 2),
                                            (1, 5, 10, 512, 2), (700, 5, 10, 512, 2)])
@pytest.mark.parametrize("prec", ["tf32x3", "f16x3", "bf16"])
def test_rank_loss_fused_operand_only_form(B, C, Nn, N, norm, prec):
    """The trainer's form of K2+K3 -- operand-only output, no score blobs -- runs the second-generation kernel (packed
    fp32 pairs, output scales folded into the per-row coefficients, reciprocal-based scalar chain): against the float64
    reference and the two-kernel path.  (700 items: more than one item per CTA and several CTA groups.)"""
    R = C + Nn
    g = torch.Generator(device="cuda").manual_seed(B + N)
    H = torch.relu(torch.randn(R * B, N, device="cuda", generator=g))
    H = (H * (torch.rand(R * B, N, device="cuda", generator=g) < 0.3) * 3.0).contiguous()
    H[min(5, R * B - 1)] = 0.0
    cfg = ops.rank_cfg(B, C, Nn, N, margin=2.0, norm=norm)
    a = ops.rank_loss_forward(H, cfg)
    b, dZ_none, op, db = ops.rank_loss_fused(H, cfg, 0.7, True, 10.0, prec=prec, want_dz=False, want_scores=False)
    assert dZ_none is None
    loss, viol, st, sn, dH_ref, dZ_ref = _rank_ref64(H, B, C, Nn, 2.0, norm, 0.7, 10.0)
    assert rel(b["stats"], a["stats"]) < 1e-6
    assert torch.equal(a["item_viol"], b["item_viol"]) and torch.equal(a["violations"], b["violations"])
    assert abs(b["loss"].item() - loss) < 1e-5 * max(1, abs(loss)) and b["violations"].item() == viol
    if prec == "f16x3":
        got, tol = op.dequant(), 1e-5
    elif prec == "bf16":
        got, tol = op.hi.float(), 1e-2
    else:
        planes = op.lo.view(-1).view(torch.bfloat16).view(2, R * B, N)   # lo = planes [bf16(x) | bf16(x - hi)], hi = tf32(x)
        got, tol = op.hi + planes[1].float(), 1e-5
    assert rel(got, dZ_ref) < tol, rel(got, dZ_ref)
    assert rel(db, dZ_ref.sum(0)) < 1e-5


def test_rank_loss_fused_unsupported_shapes(monkeypatch):
    assert ops.rank_loss_fused_supported(ops.rank_cfg(8, 5, 30, 512))         # R > 32: the wide (two-phase) kernel
    assert not ops.rank_loss_fused_supported(ops.rank_cfg(8, 5, 10, 2048))    # N > 1024
    assert not ops.rank_loss_fused_supported(ops.rank_cfg(8, 5, 256, 512))    # more negatives than the wide kernel's scalar stage takes


@pytest.mark.parametrize("B,C,Nn,N,norm", [(37, 17, 50, 1024, 2), (9, 9, 40, 512, 1), (5, 3, 33, 100, 2), (700, 5, 30, 256, 2),
                                            (1, 17, 50, 1024, 2)])
@pytest.mark.parametrize("prec", ["fp32_simt", "tf32x3", "f16x3", "bf16"])
def test_rank_loss_wide_equals_two_kernel_path(B, C, Nn, N, norm, prec):
    """Items with more than 32 rows (the large-window configuration: 16 context shots + 50 negatives): one launch, two
    phases per item, the second reading the rows back from L2.  Same formulas, trees and orders as the forward +
    backward kernels: everything but the column sums (db, fixed order here, atomics there) must be bit-identical."""
    R = C + Nn
    g = torch.Generator(device="cuda").manual_seed(B + N)
    H = torch.relu(torch.randn(R * B, N, device="cuda", generator=g))
    H = (H * (torch.rand(R * B, N, device="cuda", generator=g) < 0.3) * 3.0).contiguous()
    H[min(5, R * B - 1)] = 0.0
    cfg = ops.rank_cfg(B, C, Nn, N, margin=2.0, norm=norm)
    assert ops.rank_loss_fused_supported(cfg)
    a = ops.rank_loss_forward(H, cfg)
    dZ_a, op_a, db_a = ops.rank_loss_backward(H, cfg, a["stats"], 0.7, True, 10.0, prec=prec)
    b, dZ_b, op_b, db_b = ops.rank_loss_fused(H, cfg, 0.7, True, 10.0, prec=prec)
    for k in ("stats", "target_score", "neg_score", "item_viol", "violations", "item_loss"):
        assert torch.equal(a[k], b[k]), k
    assert rel(b["loss"], a["loss"]) < 1e-6
    assert torch.equal(dZ_a, dZ_b)
    assert rel(db_b, db_a) < 1e-5
    if prec != "fp32_simt":
        assert torch.equal(op_a.hi, op_b.hi) and (op_a.lo is None or torch.equal(op_a.lo, op_b.lo))
    loss, viol, st, sn, dH_ref, dZ_ref = _rank_ref64(H, B, C, Nn, 2.0, norm, 0.7, 10.0)
    assert abs(b["loss"].item() - loss) < 1e-5 * max(1, abs(loss)) and b["violations"].item() == viol
    assert rel(dZ_b, dZ_ref) < 1e-5
    assert rel(db_b, dZ_ref.sum(0)) < 1e-5
    # operand-only form (what the trainer asks for; there the column sums go through a workspace in a fixed order --
    # tests/test_gpu_fullsize.py checks that they repeat bit for bit)
    c1, _, op_c, db_c1 = ops.rank_loss_fused(H, cfg, 0.7, True, 10.0, prec=prec, want_dz=(prec == "fp32_simt"), want_scores=False)
    assert torch.equal(c1["loss"], b["loss"]) and rel(db_c1, db_b) < 1e-5
    if prec != "fp32_simt":
        assert torch.equal(op_c.hi, op_b.hi)


@pytest.mark.parametrize("prec", ["tf32x3", "f16x3", "bf16"])
def test_rank_loss_backward_operand_copies(prec):
    B, C, Nn, N = 32, 5, 10, 512
    H = torch.relu(torch.randn((C + Nn) * B, N, device="cuda")).contiguous()
    cfg = ops.rank_cfg(B, C, Nn, N)
    out = ops.rank_loss_forward(H, cfg)
    dZ, op, _ = ops.rank_loss_backward(H, cfg, out["stats"], 1.0, True, 10.0, prec=prec)
    if prec == "bf16":
        assert torch.equal(op.hi, dZ.to(torch.bfloat16))
    elif prec == "f16x3":
        check_f16x3_operand(op, dZ)
    else:
        check_x3_operand(op, dZ)


# ------------------------------------------------------------------------------------------------
# K4 update
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("reg_type", [2, 1])
@pytest.mark.parametrize("count", [512 * 4096, 515])
def test_sgd_update_matches_oracle(oracle, reg_type, count):
    rng = np.random.RandomState(count)
    W = rng.normal(0, 0.01, count).astype(np.float32)
    g = rng.normal(0, 0.1, (3, count)).astype(np.float32)
    h = rng.normal(0, 0.001, count).astype(np.float32)
    Wd, hd, gd = cuda(W), cuda(h), cuda(g)
    diff = torch.empty_like(Wd)
    ops.sgd_update(Wd, gd, hd, 2e-3, 0.9, 5e-4, reg_type=reg_type, grad_scale=0.5, diff_out=diff)
    Wr, dr, hr = oracle.sgd_update(W, 0.5 * g.sum(0), h, 2e-3, 0.9, 5e-4, reg_type)
    assert rel(Wd, Wr) < 1e-6 and rel(hd, hr) < 1e-6 and rel(diff, dr) < 1e-6


@pytest.mark.parametrize("prec", ["tf32x3", "f16x3", "bf16"])
def test_sgd_update_refreshes_operand_copies(prec):
    count = 64 * 128
    W = torch.randn(count, device="cuda") * 0.01
    g = torch.randn(count, device="cuda")
    h = torch.zeros(count, device="cuda")
    Wop = ops.prepare_operand(W, prec)                 # f16x3: also records max|W|, from which the update rescales
    ops.sgd_update(W, g, h, 1e-2, 0.9, 0.0, prec=prec, Wop=Wop)
    if prec == "bf16":
        assert torch.equal(Wop.hi, W.to(torch.bfloat16))
    elif prec == "f16x3":
        check_f16x3_operand(Wop, W)
    else:
        check_x3_operand(Wop, W)


def test_f16x3_operand_scale_tracks_and_saturates():
    """The header protocol: prepare measures, rescale maps max|x| to ~2^target, a stale scale saturates instead of
    overflowing, and the next rescale recovers."""
    x = torch.randn(4096, device="cuda") * 3e-4
    op = ops.prepare_operand(x, "f16x3")
    check_f16x3_operand(op, x)
    assert 2 ** 9 <= float(x.abs().max()) * op.scale < 2 ** 10
    # the tensor grows 1000x under the old scale: conversions clamp to +-65504 (no inf), the maximum is recorded
    big = x * 1000.0
    W = big.clone(); h = torch.zeros_like(W); g = torch.zeros_like(W)
    s_old = op.scale
    from videovector_b200._lib import check
    check(ops._lib.load().vv_sgd_update(ops._ptr(W), ops._ptr(g), 1, W.numel(), ops._ptr(h), None, W.numel(), 0.0, 0.0, 0.0, 2, 1.0,
                                        ops._ptr(op.hi), ops._ptr(op.lo), ops.PREC["f16x3"], ops._stream()))
    assert op.scale == s_old and bool(torch.isfinite(op.hi.float()).all()) and float(op.hi.float().abs().max()) == 65504.0
    assert op.absmax == float(big.abs().max().item())
    ops.operand_rescale(op, "f16x3")
    check(ops._lib.load().vv_sgd_update(ops._ptr(W), ops._ptr(g), 1, W.numel(), ops._ptr(h), None, W.numel(), 0.0, 0.0, 0.0, 2, 1.0,
                                        ops._ptr(op.hi), ops._ptr(op.lo), ops.PREC["f16x3"], ops._stream()))
    check_f16x3_operand(op, big)


# ------------------------------------------------------------------------------------------------
# standalone layer kernels vs the oracle's layer restatements
# ------------------------------------------------------------------------------------------------
def test_standalone_layers_match_oracle(vvlib, oracle):
    from videovector_b200.ops import _ptr, _stream
    from videovector_b200._lib import check
    import ctypes as C
    rng = np.random.RandomState(1)
    x = rng.normal(0, 1, (33, 77)).astype(np.float32); dy = rng.normal(0, 1, (33, 77)).astype(np.float32)
    xd, dyd = cuda(x), cuda(dy)
    y = torch.empty_like(xd)
    check(vvlib.vv_relu_forward(_ptr(xd), xd.numel(), 0.1, _ptr(y), _stream()))
    assert np.array_equal(y.cpu().numpy(), oracle.relu_forward(x, 0.1))
    check(vvlib.vv_relu_backward(_ptr(xd), _ptr(dyd), xd.numel(), 0.1, _ptr(y), _stream()))
    assert np.allclose(y.cpu().numpy(), oracle.relu_backward(x, dy, 0.1), rtol=1e-7)
    check(vvlib.vv_l2norm_forward(_ptr(xd), 33, 77, _ptr(y), _stream()))
    assert rel(y, oracle.normalization_forward(x)) < 1e-6
    check(vvlib.vv_l2norm_backward(_ptr(xd), _ptr(dyd), 33, 77, _ptr(y), _stream()))
    assert rel(y, oracle.normalization_backward(x, dy)) < 1e-5
    s = torch.empty((33, 10), device="cuda")
    check(vvlib.vv_rowsum_forward(_ptr(xd), 33, 77, 10, _ptr(s), _stream()))
    assert rel(s, oracle.sum_forward(x, 10)) < 1e-6
    ds = cuda(rng.normal(0, 1, (33, 10)).astype(np.float32))
    check(vvlib.vv_rowsum_backward(_ptr(ds), 33, 77, 10, _ptr(y), _stream()))
    assert rel(y, oracle.sum_backward(ds.cpu().numpy(), 77)) < 1e-6
    m = (rng.uniform(0, 1, x.shape) < 0.5).astype(np.uint32)
    md = cuda(m.astype(np.int32))
    check(vvlib.vv_dropout_forward(_ptr(xd), _ptr(md), DROPOUT_MASK01, xd.numel(), 0.5, _ptr(y), _stream()))
    assert np.array_equal(y.cpu().numpy(), oracle.dropout_forward(x, m, 0.5))
    t = cuda(rng.normal(0, 10, (10, 5)).astype(np.float32)); b = cuda(rng.normal(0, 10, (10, 5)).astype(np.float32))
    loss = torch.zeros(1, device="cuda"); viol = torch.zeros(1, device="cuda"); hinge = torch.zeros(50, device="cuda")
    for norm in (1, 2):
        check(vvlib.vv_max_margin_forward(_ptr(t), _ptr(b), 50, 1.0, norm, _ptr(hinge), _ptr(loss), _ptr(viol), _stream()))
        lr, vr, _ = oracle.max_margin_forward(t.cpu().numpy(), b.cpu().numpy(), 1.0, norm)
        assert abs(loss.item() - lr) < 1e-5 * max(1, abs(lr)) and viol.item() == vr
        dt = torch.zeros_like(t); dbg = torch.zeros_like(t)
        check(vvlib.vv_max_margin_backward(_ptr(t), _ptr(b), 50, 1.0, norm, 1.0, _ptr(dt), _ptr(dbg), _stream()))
        rt, rb = oracle.max_margin_backward(t.cpu().numpy(), b.cpu().numpy(), 1.0, norm, 1.0)
        assert rel(dt, rt) < 1e-6 and rel(dbg, rb) < 1e-6
    bots = [cuda(rng.normal(0, 1, (5, 9)).astype(np.float32)) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[t_.data_ptr() for t_ in bots]); co = (C.c_float * 4)(0.25, 0.25, 0.25, 0.25)
    top = torch.empty_like(bots[0])
    check(vvlib.vv_eltwise_sum_forward(ptrs, co, 4, 45, _ptr(top), _stream()))
    assert rel(top, oracle.eltwise_sum_forward([t_.cpu().numpy() for t_ in bots], [0.25] * 4)) < 1e-6


def test_errors_are_loud(vvlib):
    """Bad arguments return an error code and a message; nothing falls back to the CPU."""
    X = torch.zeros((8, 12), device="cuda")
    with pytest.raises(Exception):
        ops.ip_forward(ops.OperandT(X), ops.OperandT(X), None, 8, 8, 12, "bf16")     # K % 8 != 0 on the TC path
    assert "K" in vvlib.vv_last_error().decode() or "tensor-core" in vvlib.vv_last_error().decode()
    with pytest.raises(Exception):
        ops.rank_loss_forward(torch.zeros((15, 6), device="cuda"), ops.rank_cfg(1, 5, 10, 6))   # N % 4 != 0


def test_bank_operand_row_interleaved_layout(vvlib):
    """vv_prepare_bank_operand (F16X3, K % 64 == 0): the same fp16 pairs as the two-plane form, stored per row as blocks
    [64 x h0 | 64 x h1]; header layout flag 1; other K / precisions fall back to the plain operand."""
    from videovector_b200.ops import _ptr, _stream
    rows, K = 37, 256
    bank = torch.relu(torch.randn(rows, K, device="cuda")) * 3
    planar = ops.prepare_operand(bank, "f16x3")
    il = ops.alloc_operand((rows, K), "f16x3")
    ops.check(vvlib.vv_prepare_bank_operand(_ptr(bank), rows, K, ops.PREC["f16x3"], _ptr(il.hi), _ptr(il.lo), _stream()))
    hdr = il.block[:16].view(torch.int32)
    assert int(hdr[3]) == 1 and int(planar.block[:16].view(torch.int32)[3]) == 0
    raw = il.block[128:128 + rows * K * 4].view(torch.float16).view(rows, K // 64, 2, 64)
    s_il, s_pl = il.scale, planar.scale                              # bank copy: max -> 2^12, plain operand: 2^10
    assert s_il == 4 * s_pl
    assert torch.equal(raw[:, :, 0, :].reshape(rows, K), (bank * s_il).to(torch.float16))
    assert torch.equal(raw[:, :, 1, :].reshape(rows, K), (bank * s_il - (bank * s_il).to(torch.float16).float()).to(torch.float16))
    # K not a multiple of 64 -> two planes
    bank2 = torch.relu(torch.randn(5, 72, device="cuda"))
    op2 = ops.alloc_operand((5, 72), "f16x3")
    ops.check(vvlib.vv_prepare_bank_operand(_ptr(bank2), 5, 72, ops.PREC["f16x3"], _ptr(op2.hi), _ptr(op2.lo), _stream()))
    assert int(op2.block[:16].view(torch.int32)[3]) == 0 and torch.equal(op2.hi, (bank2 * op2.scale).to(torch.float16))

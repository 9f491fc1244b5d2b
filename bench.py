#!/usr/bin/env python
"""bench.py -- training triplets/sec of the temporal-context embedding step on B200.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W    (the reference's CPU path, rank 0 only)

Prints ONE JSON line (rank 0); its "configs" object carries the other BASELINE configurations measured in the same run
(bf16_dp = configs[2], large_window = configs[3], inference = configs[4]; --no-extra-configs skips them).
Headline workload (BASELINE.json configs[1] per GPU, weak scaling):
synthetic post-ReLU 4096-d features resident in HBM, 512-d embedding, window +-2 (C=5),
10 negatives, B = 4096 triplets per GPU per step, dropout 0.9 (Philox), squared hinge margin 2,
SGD momentum 0.9 / decay 5e-4 / inv LR policy.  Default precision: f16x3 (scaled fp16 split operands, three
tensor-core products, fp32 accumulate = the fp32-parity mode config[1] names; tf32x3 is the other parity mode);
--precision bf16|tf32 for config[2]'s mode.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(C=5, Nn=10, K=4096, N=512, B=4096, V=8192, S=32, P=5000, swap=50, max_same=6)
CPU_SAMPLE_B = 256          # bounded CPU sample: items per oracle step (same K, N, C, Nn)
METRIC = "training triplets/sec"
CPU_NOTE = {
    "reference": "the reference's own sources compiled unmodified against shim headers (oracle/_ref): its "
                 "VideoSampledShotsDataLayer (prefetch thread, in-memory fake LMDB), Net (net.cpp) and SGDSolver (solver.cpp) "
                 "in Caffe CPU mode, OpenBLAS from the SciPy wheel; one step = one iteration of Solver::Solve's loop",
    "port": "oracle/vv_oracle.cpp = restated reference CPU layers + solver + sampler, OpenBLAS from the SciPy wheel "
            "(oracle/_ref was not built on this machine)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sus=float(d["bf16_tflops_sustained"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples drawing more than half of the peak power seen
        pmax = max(power)
        load = [s for s, p in zip(sm, power) if p >= 0.5 * pmax] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "power_w_max": pmax, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (port of the reference's CPU layers) on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, threads=0):
    """One 'step' = one training iteration on a bounded sample of CPU_SAMPLE_B items of the workload, on the host cores.
    With oracle/_ref (the reference's own data layer, Net and SGDSolver sources, shim-compiled) it is one iteration of the
    reference's Solver::Solve loop -- kind "reference"; without it the oracle port runs sampler + net + update.
    Returns (triplets/s, ms/step, cores, phase dict, kind)."""
    from oracle import pyoracle as orc
    from oracle import pyref
    from videovector_b200 import ops
    c = CFG
    B = CPU_SAMPLE_B
    V, S = 512, 16                                       # 8192 shots >= the 5000-entry negative buffer
    if threads <= 0:                                     # every host core this process may use, whatever OMP_NUM_THREADS says
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = orc.use_openblas(threads)                    # same OpenBLAS instance the reference library links
    use_ref = pyref.available() and hasattr(pyref.lib(), "ref_solver_create")
    feat = ops.bank_host(V * S, c["K"], 1234)
    vid, off, sid = ops.synthetic_videos(V, S)
    if use_ref:
        # the reference's whole pipeline: data layer (its own prefetch thread) + Net + SGDSolver, shipped hyper-parameters
        rng = np.random.RandomState(1701)
        W = rng.normal(0, 0.001, (c["N"], c["K"])).astype(np.float32)
        sol = pyref.Solver(vid, off, sid, feat, W, np.zeros(c["N"], np.float32), B, c["C"], c["Nn"], c["P"], c["swap"],
                           c["max_same"], margin=2.0, norm=2, base_lr=1e-3, momentum=0.9, weight_decay=5e-4, lr_policy="inv",
                           gamma=1e-3, power=0.75, dropout_ratio=0.9)
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            loss, _ = sol.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        sol.close()
        orc.use_builtin_blas()
        ms = 1e3 * float(np.mean(times))
        return B / (ms * 1e-3), ms, cores, {"solver_step": ms, "last_loss": loss}, "reference"
    smp = orc.Sampler(vid, off, sid, feat, c["K"], B, c["C"], c["Nn"], c["P"], c["swap"], c["max_same"], 100, seed=1)   # port
    rng = np.random.RandomState(1701)
    W = rng.normal(0, 0.001, (c["N"], c["K"])).astype(np.float32)
    b = np.zeros(c["N"], np.float32)
    hW = np.zeros_like(W); hb = np.zeros_like(b)
    R = c["C"] + c["Nn"]
    mrng = np.random.RandomState(7)
    times, phases = [], np.zeros(8)
    for it in range(warmup + steps):
        mask = (mrng.uniform(0, 1, (R * B, c["N"])) > 0.9).astype(np.uint32)   # mask generation not timed
        t0 = time.perf_counter()
        idx, quirk, data = smp.next()
        t1 = time.perf_counter()
        out = orc.net_forward_backward(data, W, b, mask, B, c["C"], c["Nn"], margin=2.0, norm=2, dropout_ratio=0.9,
                                       want=("loss", "violations", "dW", "db"))
        t2 = time.perf_counter()
        rate = orc.learning_rate("inv", 1e-3, 1e-3, 0.75, 1, it)
        W, _, hW = orc.sgd_update(W, out["dW"], hW, rate, 0.9, 5e-4)
        b, _, hb = orc.sgd_update(b, out["db"], hb, rate * 2, 0.9, 0.0)
        t3 = time.perf_counter()
        if it >= warmup:
            times.append(t3 - t0)
            phases[:5] += out["phase_seconds"][:5]
            phases[5] += t1 - t0; phases[6] += t3 - t2
    smp.close()
    orc.use_builtin_blas()
    ms = 1e3 * float(np.mean(times))
    names = ["slice_concat", "forward", "loss_forward", "backward", "fc7_backward", "sampler", "sgd_update"]
    ph = {k: 1e3 * phases[i] / len(times) for i, k in enumerate(names) if phases[i] > 0}
    return B / (ms * 1e-3), ms, cores, ph, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, ms, cores, ph, kind = cpu_reference_run(args.steps, max(args.warmup, 1))
    sample = "B=%d items per step (1/%d of the %d-item GPU step), same K/N/C/Nn; %d timed steps" % (
        CPU_SAMPLE_B, CFG["B"] // CPU_SAMPLE_B, CFG["B"], args.steps)
    c = CFG
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "triplets/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # what THIS arm steps: the same net on the host cores, on a bounded sample of the GPU arm's batch
        "config": {"workload": "videovec_embedding context-ranking training step in Caffe CPU mode (BASELINE configs[0]/[1] net): "
                               "4096-d features -> %d-d embedding, window +-%d, %d negatives, dropout 0.9, squared hinge margin 2, "
                               "SGD momentum; bounded sample of the GPU arm's B=%d step" % (c["N"], c["C"] // 2, c["Nn"], c["B"]),
                   "global_batch": CPU_SAMPLE_B, "K": c["K"], "N": c["N"], "C": c["C"], "Nn": c["Nn"],
                   "parallelism": "cpu (%d host threads)" % cores, "gpu_arm_global_batch": c["B"] * max(args.gpus, 1)},
        "cpu_baseline": {"value": val, "unit": "triplets/s", "cores": cores, "kind": kind, "sample": sample,
                         "phase_ms": ph, "note": CPU_NOTE[kind]},
        "e2e": {"value": val, "unit": "triplets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# bytes per element of a GEMM operand copy, and tensor-pipe work per algorithmic product in units of one
# full-rate 16-bit MMA (tf32 runs at half rate: tf32x3 = tf32 hi*hi (2) + two bf16 cross terms; f16x3 = three fp16)
OPERAND_BYTES = {"tf32x3": 8, "f16x3": 4, "bf16": 2, "tf32": 4, "fp32_simt": 4}
MMA_UNITS = {"tf32x3": 4, "f16x3": 3, "bf16": 1, "tf32": 2, "fp32_simt": 0}
DTYPE = {"tf32x3": "f32 (tf32 + bf16 split operands, 3 tensor-core products, fp32 accumulate)",
         "f16x3": "f32 (scaled fp16 split operands, 3 tensor-core products, fp32 accumulate)",
         "bf16": "bf16", "tf32": "tf32", "fp32_simt": "f32"}


def fused_gather_for(prec, args):
    """K0 folded into the GEMM producers: the default for the 2-byte operand formats; --materialised-gather turns it off."""
    return prec in ("f16x3", "bf16") and not getattr(args, "materialised_gather", False)


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def sampler_streams(args, world):
    """Reference-exact sampler streams per rank (each on its own sub-shard of the rank's videos, each with a native prefetch
    thread).  One stream draws ~3.3 M items/s on one host core; default = what the rank's share of the host cores allows."""
    if args.sampler_streams > 0:
        return args.sampler_streams
    per_rank = max(1, host_cores() // max(world, 1))
    return max(1, min(4, per_rank - 1))          # one core stays with the rank's main thread


def workload_config(args, world, c=None, prec=None, name="BASELINE configs[1] per GPU", streams=None):
    c = c or CFG
    prec = prec or getattr(args, "precision", "f16x3")
    smode = getattr(args, "sampler", "sharded")
    return {"workload": "videovec_embedding context-ranking training step (%s): "
                        "4096-d features -> %d-d embedding, window +-%d, %d negatives, B=%d triplets/GPU/step, "
                        "dropout 0.9, squared hinge margin 2, SGD momentum" % (name, c["N"], c["C"] // 2, c["Nn"], c["B"]),
            "global_batch": c["B"] * world, "K": c["K"], "N": c["N"], "C": c["C"], "Nn": c["Nn"],
            "parallelism": "dp%d" % world, "precision": prec,
            "dgrad": False, "bank_rows": c["V"] * c["S"],
            "sampler": smode,
            "sampler_note": ("sharded: every rank owns a shard of the videos (its own resident bank) and draws its own reference-exact "
                             "stream(s) over it -- the G-GPU index stream is NOT the 1-GPU stream; "
                             if smode == "sharded" else
                             "global: ONE reference-exact stream of G*B items per step, drawn identically on every rank, rank r takes items "
                             "[r*B, (r+1)*B) -- the G-GPU index stream is the 1-GPU stream of the global batch (SURVEY 8e); ") +
                            "%s stream(s) per rank, batches round-robin" % (streams if streams is not None else "?"),
            "gather": "fused into the GEMMs (cp.async row gather of the bank's operand copy by two producer warps)"
                      if fused_gather_for(prec, args) else "materialised X (K0 kernel)",
            "gemm": ("tcgen05.mma cta_group::2: one MMA per CTA pair, each CTA stages its own A rows and half of B"
                     if (prec == "f16x3" and fused_gather_for(prec, args) and os.environ.get("VV_GEMM_2CTA", "1") != "0")
                     else "tcgen05.mma cta_group::1, 2-CTA clusters with B multicast") if prec != "fp32_simt" else "fp32 SIMT",
            "l2": "inputs larger than L2: every GEMM streams %.2f GB of gathered operand rows (> 126 MB L2)" % (
                (c["C"] + c["Nn"]) * c["B"] * c["K"] * OPERAND_BYTES[prec] / 1e9)}


class GlobalSliceSampler:
    """SURVEY 8e's partition: one sampler stream of world*B items per step, identical on every rank; this rank's slice."""

    def __init__(self, smp, rank, world, B):
        self.smp, self.rank, self.world, self.B = smp, rank, world, B
        R = smp.R
        self._i = np.empty((world * B, R), np.int32); self._q = np.empty((world * B, R), np.int32)

    def next_into(self, idx, quirk):
        self.smp.next_into(self._i, self._q)
        idx[...] = self._i[self.rank * self.B:(self.rank + 1) * self.B]
        quirk[...] = self._q[self.rank * self.B:(self.rank + 1) * self.B]

    def prefetch(self, depth):
        self.smp.prefetch(depth)

    @property
    def ready(self):
        return self.smp.ready

    def close(self):
        self.smp.close()


# ---------------------------------------------------------------------------------------------------
# one training configuration on this rank: `value` leg, per-kernel leg, `e2e` leg
# ---------------------------------------------------------------------------------------------------
def run_training(c, prec, args, steps, warmup, world, rank, stream, pk, want_e2e=True):
    import torch
    import torch.distributed as dist
    from videovector_b200 import ops
    from videovector_b200._lib import DROPOUT_HASH
    B, C, Nn, K, N = c["B"], c["C"], c["Nn"], c["K"], c["N"]
    R = C + Nn
    fg = fused_gather_for(prec, args)
    nstreams = sampler_streams(args, world)
    # synthetic feature bank (resident in HBM) and the host sampler
    vid, off, sid = ops.synthetic_videos(c["V"], c["S"])
    if args.sampler == "global":
        bank = ops.fill_bank(c["V"] * c["S"], K, 1234)
        smp = GlobalSliceSampler(ops.Sampler(vid, off, sid, B * world, C, Nn, c["P"], c["swap"], c["max_same"], 100, rand_seed=1),
                                 rank, world, B)
        nstreams = 1
    else:
        bank = ops.fill_bank(c["V"] * c["S"], K, 1234 + rank)
        smp = ops.MultiSampler(vid + rank * c["V"], off, sid, B, C, Nn, c["P"], c["swap"], c["max_same"], 100,
                               rand_seed=1 + 16 * rank, streams=nstreams)
    out = {"streams": nstreams}
    with torch.cuda.stream(stream):
        tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, prec=prec, dropout_ratio=0.9, dropout_mode=DROPOUT_HASH,
                                         dropout_seed=7, world_size=world, rank=rank), stream=stream)
        g = torch.Generator(device="cuda").manual_seed(1701)
        W0 = torch.randn(N, K, device="cuda", generator=g) * 0.001          # gaussian filler std 0.001, bias 0
        tr.set_weights(W0, torch.zeros(N, device="cuda"))
        if fg:
            tr.set_bank(bank)          # one-time operand copy of the resident bank; the GEMMs then gather its rows themselves
        if world > 1:
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(ops.dp_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            tr.dp_init(bytes(idt.cpu().numpy().tobytes()))
        out["dp_mode"] = tr.dp_mode + ((" (" + tr.dp_mode_reason + ")") if tr.dp_mode_reason else "")
        total = warmup + steps
        # ---- leg 1: `value` -- inputs already resident in HBM when the timed region starts
        idx_host = torch.empty((total, B, R), dtype=torch.int32).pin_memory()
        qk_host = torch.empty((total, B, R), dtype=torch.int32).pin_memory()
        for i in range(total):
            smp.next_into(idx_host[i].numpy(), qk_host[i].numpy())
        idx_dev = idx_host.cuda(non_blocking=True); qk_dev = qk_host.cuda(non_blocking=True)
        for i in range(warmup):
            tr.step(bank, idx_dev[i], qk_dev[i], None, it=i)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(warmup, total):
            tr.step(bank, idx_dev[i], qk_dev[i], None, it=i)
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        out["ms_total"] = e0.elapsed_time(e1)
        out["launches"] = tr.last_launches * steps
        out["loss"] = float(tr.tensor("loss").item())

        # ---- leg 2: per-kernel timing inside the step (CUDA events on the launching stream)
        tr.set_timing(True)
        nt = min(steps, 20)
        for i in range(nt):
            tr.step(bank, idx_dev[warmup + i % steps], qk_dev[warmup + i % steps], None, it=total + i)
        out["phase"], _ = tr.phase_ms()
        tr.set_timing(False)
        out["nsplit"] = tr._lib.vv_ip_wgrad_auto_nsplit(R * B, N, K, ops.PREC[prec])
        # how evenly the ranks run: every rank's own compute per step (everything before the exchange).  In a synchronous
        # step the exchange ends when the slowest rank arrives, so the spread shows up as waiting inside the exchange kernel
        comp = sum(out["phase"][k] for k in ("gather", "fc7_forward", "rank_loss_forward", "rank_loss_backward", "wgrad"))
        if world > 1:
            allc = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
            dist.all_gather(allc, torch.tensor([comp], device="cuda", dtype=torch.float64))
            out["rank_compute_ms"] = [float(x.item()) for x in allc]
        else:
            out["rank_compute_ms"] = [comp]

        # ---- leg 3: `e2e` -- the public API with HOST buffers: sampler (prefetch threads, as the reference's
        # prefetching data layer) -> pinned host indices -> H2D -> step -> D2H loss, all inside the timed region
        if want_e2e:
            nbuf = 8                  # batches the sampler threads may run ahead (absorbs host jitter; the average rates decide)
            hi = [torch.empty((B, R), dtype=torch.int32).pin_memory() for _ in range(nbuf)]
            hq = [torch.empty((B, R), dtype=torch.int32).pin_memory() for _ in range(nbuf)]
            di = [torch.empty((B, R), dtype=torch.int32, device="cuda") for _ in range(nbuf)]
            dq = [torch.empty((B, R), dtype=torch.int32, device="cuda") for _ in range(nbuf)]
            loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
            smp.prefetch(nbuf)        # native prefetch thread(s) (vv_sampler_prefetch), the reference's InternalThread
            evs = [torch.cuda.Event() for _ in range(nbuf)]         # H2D of buffer k complete
            used = [torch.cuda.Event() for _ in range(nbuf)]        # the step that read device buffer k complete
            cstream = torch.cuda.Stream()                           # copy stream: step i+1's indices arrive under step i

            def e2e_step(i):
                k = i % nbuf
                if i >= nbuf:
                    evs[k].synchronize()          # the H2D copy that last read this pinned buffer (nbuf steps ago) is done
                    cstream.wait_event(used[k])   # ... and the step that consumed device buffer k has finished with it
                smp.next_into(hi[k].numpy(), hq[k].numpy())
                with torch.cuda.stream(cstream):
                    di[k].copy_(hi[k], non_blocking=True); dq[k].copy_(hq[k], non_blocking=True)
                    evs[k].record(cstream)
                stream.wait_event(evs[k])
                tr.step(bank, di[k], dq[k], None, it=2 * total + i)
                used[k].record(stream)
                loss_host.copy_(tr.tensor("db_raw_ext")[N:N + 2], non_blocking=True)      # loss + violations, 8 bytes D2H
            for i in range(warmup):
                e2e_step(i)
            stream.synchronize()
            if world > 1:
                dist.barrier()
            ready0 = smp.ready        # batches drawn ahead when the clock starts ...
            t0 = time.perf_counter()
            for i in range(warmup, warmup + steps):
                e2e_step(i)
            stream.synchronize()
            # ... must be drawn ahead again when it stops: the sampler work of exactly `steps` batches lies inside the region
            # (a sampler-bound run would otherwise borrow up to nbuf batches drawn before t0)
            while smp.ready < ready0 and time.perf_counter() - t0 < 120.0:
                time.sleep(20e-6)
            t1 = time.perf_counter()
            smp.prefetch(0)
            out["e2e_ms"] = 1e3 * (t1 - t0)
            out["e2e_ring"] = {"ahead_at_t0": ready0, "ahead_at_t1": smp.ready, "depth": nbuf}
        tr.close()
    smp.close()
    del bank
    torch.cuda.empty_cache()
    return out


def training_kernels(c, prec, out, args, pk):
    """Per-kernel roofline table of one training configuration from the in-step CUDA-event timings."""
    B, C, Nn, K, N = c["B"], c["C"], c["Nn"], c["K"], c["N"]
    R = C + Nn; M = R * B
    phase = out["phase"]
    flops = 2.0 * M * N * K
    units, opb = MMA_UNITS[prec], OPERAND_BYTES[prec]
    opw = opb if prec not in ("tf32", "fp32_simt") else 0          # bytes of a separate operand copy per element
    tensor_peak = pk["bf16_sus"]       # the kernels are timed inside a long step: sustained peak
    kern = {}
    for name in ("fc7_forward", "wgrad"):
        ms = phase[name]
        tf = flops / (ms * 1e-3) / 1e12 if ms > 0 else None
        kern[name] = {"ms": ms, "bound": "tensor", "achieved_tflops": tf, "frac": tf / tensor_peak if tf else None,
                      "tensor_pipe_frac": tf * units / tensor_peak if tf else None}
    fused_rank = phase["rank_loss_forward"] == 0
    p2p = str(out.get("dp_mode", "")).startswith("p2p")
    fg = fused_gather_for(prec, args)
    bytes_alg = {"rank_loss_forward": M * N * 4,
                 # fused K2+K3 reads H once; the two-kernel K3 reads it again
                 "rank_loss_backward": M * N * (4 + (opw if opw else 4)),
                 # read W, dW, hist + write W, hist = 5*N*K*4, + the refreshed operand copy of W
                 "sgd_update": N * K * (5 * 4 + opw)}
    if not fg:
        bytes_alg["gather"] = M * K * (4 + (opw if opw else 4))
    for name, by in bytes_alg.items():
        ms = phase[name]
        if (name == "rank_loss_forward" and fused_rank) or ms <= 0:
            continue
        key = "rank_loss_fused" if (name == "rank_loss_backward" and fused_rank) else name
        kern[key] = {"ms": ms, "bound": "hbm", "achieved_gbs": by / (ms * 1e-3) / 1e9, "frac": by / (ms * 1e-3) / 1e9 / pk["hbm"]}
    if "sgd_update" in kern:
        kern["sgd_update"]["note"] = ("algorithmic bytes = 5*N*K*4 + operand refresh; the kernel also reads the %d split-K slabs the wgrad "
                                      "wrote (not counted as algorithmic)" % out["nsplit"])
        if p2p:
            kern["dp_exchange_update"] = kern.pop("sgd_update")
            kern["dp_exchange_update"].update({
                "bound": "nvlink + hbm (not graded)", "frac": None,
                "note": "ONE kernel: split-K sum -> rows pushed to their owner rank over NVLink -> owner adds the G contributions, "
                        "updates its N/G rows -> pushes the new operand rows to every rank; includes waiting for the slowest rank",
                "nvlink_bytes_out": (out["world"] - 1) * (N * K // out["world"]) * (4 + opw)})
    if phase["sgd_update"] <= 0 or p2p:
        kern["wgrad"]["note"] = ("includes the split-K finish in the kernel's drain phase: the %d partial slabs summed in slab order, then "
                                 "%s (no separate update launch)" % (out["nsplit"], "the rows pushed to their owner ranks over NVLink" if p2p
                                                                      else "decay + momentum + update + operand refresh of W and the bias update"))
    if fg and phase["gather"] > 0:
        kern["gather_plan"] = {"ms": phase["gather"], "bound": "latency (index traffic only; not graded)", "frac": None}
    if phase.get("allreduce", 0) > 0:
        # exposed time of the gradient exchange on the main stream: split-K reduce + wait for the NCCL all-reduce of
        # dW [N,K] and (db, loss, violations) -- includes waiting for the slowest rank
        kern["allreduce"] = {"ms": phase["allreduce"], "bound": "nvlink", "bytes": N * K * 4 + (N + 2) * 4}
    return kern


def run_inference(args, world, rank, stream, pk, prec):
    """BASELINE configs[4]: relu(F W^T + b) over this rank's shard of the 10 M-row sweep (rows sharded 8 ways, no collective)."""
    import torch
    import torch.distributed as dist
    from videovector_b200 import ops
    K, N = CFG["K"], CFG["N"]
    total_rows, shards = 10_000_000, 8
    rows = total_rows // shards
    with torch.cuda.stream(stream):
        F = ops.fill_bank(rows, K, 4242 + rank)
        tr = ops.Trainer(ops.trainer_cfg(4096, 5, 10, K, N, prec=prec), stream=stream)
        g = torch.Generator(device="cuda").manual_seed(1701)
        tr.set_weights(torch.randn(N, K, device="cuda", generator=g) * 0.01, torch.zeros(N, device="cuda"))
        out = torch.empty((rows, N), dtype=torch.float32, device="cuda")
        tr.extract_into(F, out)                      # warm-up (also allocates the operand staging buffer)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        reps = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            tr.extract_into(F, out)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / reps
        launches = tr.last_launches
        # e2e: HOST rows -> H2D -> extract -> D2H embeddings, on a bounded chunk of the shard (PCIe-bound by construction)
        chunk = 65536
        Fh = torch.empty((chunk, K), dtype=torch.float32).pin_memory()
        Fh.copy_(F[:chunk])
        Oh = torch.empty((chunk, N), dtype=torch.float32).pin_memory()
        Fd = torch.empty((chunk, K), dtype=torch.float32, device="cuda"); Od = torch.empty((chunk, N), dtype=torch.float32, device="cuda")
        def e2e_once():
            Fd.copy_(Fh, non_blocking=True)
            tr.extract_into(Fd, Od)
            Oh.copy_(Od, non_blocking=True)
        e2e_once(); stream.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_once()
        stream.synchronize()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / 3
        tr.close()
    del F, out
    torch.cuda.empty_cache()
    t = torch.tensor([ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    by = rows * (K + N) * 4
    fl = 2.0 * rows * K * N
    units = MMA_UNITS[prec]
    return {
        "metric": "embedding inference rows/sec", "unit": "rows/s", "value": world * rows / (ms * 1e-3), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "dtype": DTYPE[prec], "gpu_launches": launches,
        "config": {"workload": "embedding inference sweep (BASELINE configs[4]): relu(F W^T + b), 10 M synthetic 4096-d rows -> %d-d, "
                               "rows sharded 8 ways, this run holds %d of the 8 shards (one per GPU, %d rows each), no collective; "
                               "one step = one pass over the shard" % (N, world, rows),
                   "rows_per_gpu": rows, "K": K, "N": N, "precision": prec, "parallelism": "row-sharded x%d" % world},
        "roofline": {"bound": "hbm", "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": by / (ms * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                     "algorithmic_bytes_per_row": (K + N) * 4,
                     "tensor_pipe_frac": fl / (ms * 1e-3) / 1e12 * units / pk["bf16_sus"],
                     "note": "SURVEY 8(d): at N=512 with fp32 input this sits at the ridge (AI 227 flop/B); both fractions reported"},
        "e2e": {"value": world * chunk / (e2e_ms * 1e-3), "unit": "rows/s", "h2d_bytes_per_step": chunk * K * 4,
                "d2h_bytes_per_step": chunk * N * 4, "ms_per_step": e2e_ms,
                "note": "pinned host rows -> H2D -> vv_trainer_extract -> D2H embeddings on a %d-row chunk: PCIe-bound" % chunk},
    }


# ---------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = CFG
    B, C, Nn, K, N = c["B"], c["C"], c["Nn"], c["K"], c["N"]
    R = C + Nn
    prec = args.precision
    pk = peaks()
    stream = torch.cuda.Stream()

    def reduce_max(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    # ---- the headline: BASELINE configs[1] per GPU
    clocks = ClockSampler(local); clocks.start()
    out = run_training(c, prec, args, args.steps, args.warmup, world, rank, stream, pk)
    out["world"] = world
    # one sampler over the timed `value` region, the per-kernel timing steps and the timed e2e region (all the same
    # workload): a 100 ms poll would see nothing of a short --steps run otherwise; idle samples are filtered by power
    clk = clocks.stop()
    clk["window"] = "value + per-kernel + e2e legs of the headline configuration"
    ms_total, e2e_ms = reduce_max([out["ms_total"], out["e2e_ms"]])
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    e2e_val = world * B * args.steps / (e2e_ms * 1e-3)

    line = None
    if rank == 0:
        M = R * B
        units = MMA_UNITS[prec]
        tensor_peak = pk["bf16_sus"]
        kern = training_kernels(c, prec, out, args, pk)
        phase = out["phase"]
        dom = "wgrad" if phase["wgrad"] >= phase["fc7_forward"] else "fc7_forward"
        ach = kern[dom]["achieved_tflops"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
            key = prec + ("_gathered" if fused_gather_for(prec, args) else "")
            traffic = json.load(open(tpath)).get(key, {}).get(dom)
        roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": ach / tensor_peak if ach else None, "traffic": traffic,
                    "peak_source": pk["src"] + " (cuBLAS bf16, sustained)",
                    "mma_units_per_product": units,
                    "tensor_pipe_frac": ach * units / tensor_peak if ach else None,
                    "peak_burst": pk.get("bf16"),
                    "tensor_pipe_frac_of_burst": (ach * units / pk["bf16"]) if (ach and pk.get("bf16")) else None,
                    "note": "achieved = algorithmic 2*M*N*K per launch / mean launch duration (CUDA events around the kernel "
                            "inside the step), against the measured full-rate 16-bit tensor peak.  The fp32-parity modes "
                            "spend several tensor-core products per algorithmic product (mma_units_per_product, in units "
                            "of one full-rate 16-bit MMA); tensor_pipe_frac = frac x units is the pipe's utilisation"}
        cfgd = workload_config(args, world, streams=out["streams"])
        cfgd["dp_mode"] = out["dp_mode"]
        rc = out["rank_compute_ms"]
        skew = {"compute_ms_per_rank": [round(x, 4) for x in rc], "slowest_minus_fastest_ms": max(rc) - min(rc),
                "slowest_over_rank0": max(rc) / rc[0] if rc[0] > 0 else None,
                "note": "per-rank device time of gather plan + forward + rank loss + wgrad (in-step CUDA events): a synchronous "
                        "step runs at the pace of the slowest rank (per-GPU power capping), the others wait inside the exchange"}
        line = {
            "metric": METRIC, "value": value, "unit": "triplets/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE[prec],
            "data": "synthetic", "config": cfgd,
            "clocks": clk, "gpu_launches": out["launches"],
            "e2e": {"value": e2e_val, "unit": "triplets/s", "h2d_bytes_per_step": 2 * B * R * 4, "d2h_bytes_per_step": 8,
                    "ms_per_step": e2e_ms / args.steps, "ring": out["e2e_ring"],
                    "note": "host sampler on prefetch thread(s) -> pinned int32 [B,R] indices -> H2D -> vv_trainer_step -> "
                            "D2H loss; the feature bank is resident in HBM (uploaded once, like opening the LMDB); the clock stops "
                            "only when the sampler is as far ahead as it was when the clock started"},
            "roofline": roofline, "kernels": kern, "loss": out["loss"],
            "hinge_terms_per_s": value * Nn, "rank_skew": skew,
        }

    # ---- the other BASELINE configurations, same run (each with value, ms_per_step, roofline.frac, e2e)
    extra = {}
    if not args.no_extra_configs:
        xs = max(10, min(args.steps, 60))
        # configs[2]: bf16 tensor-core mode, 4096 triplets per GPU (32 k per step on 8 GPUs)
        o2 = run_training(c, "bf16", args, xs, args.warmup, world, rank, stream, pk); o2["world"] = world
        # configs[3]: large window, N=1024, C=17 (+-8), 50 negatives: the HBM-bound loss kernels matter here
        c4 = dict(c, C=17, Nn=50, N=1024)
        o4 = run_training(c4, prec, args, max(5, xs // 4), args.warmup, world, rank, stream, pk); o4["world"] = world
        for name, cc, pp, oo, st, label in (("bf16_dp", c, "bf16", o2, xs, "BASELINE configs[2] per GPU"),
                                            ("large_window", c4, prec, o4, max(5, xs // 4), "BASELINE configs[3] per GPU")):
            mt, em = reduce_max([oo["ms_total"], oo["e2e_ms"]])
            if rank == 0:
                ms1 = mt / st
                kk = training_kernels(cc, pp, oo, args, pk)
                dom = "wgrad" if oo["phase"]["wgrad"] >= oo["phase"]["fc7_forward"] else "fc7_forward"
                Bc = cc["B"]
                cf = workload_config(args, world, cc, pp, label, streams=oo["streams"]); cf["dp_mode"] = oo["dp_mode"]
                extra[name] = {
                    "metric": METRIC, "unit": "triplets/s", "value": world * Bc / (ms1 * 1e-3), "ms_per_step": ms1, "steps": st,
                    "higher_is_better": True, "scaling": "weak", "dtype": DTYPE[pp], "gpu_launches": oo["launches"], "config": cf,
                    "roofline": {"kernel": dom, "bound": "tensor", "achieved": kk[dom]["achieved_tflops"], "peak": pk["bf16_sus"],
                                 "unit": "TFLOP/s", "frac": kk[dom]["frac"], "tensor_pipe_frac": kk[dom]["tensor_pipe_frac"],
                                 "mma_units_per_product": MMA_UNITS[pp], "traffic": None},
                    "kernels": kk, "loss": oo["loss"],
                    "e2e": {"value": world * Bc * st / (em * 1e-3), "unit": "triplets/s", "ms_per_step": em / st,
                            "h2d_bytes_per_step": 2 * Bc * (cc["C"] + cc["Nn"]) * 4, "d2h_bytes_per_step": 8, "ring": oo["e2e_ring"]}}
        # configs[4]: inference sweep, rows sharded
        inf = run_inference(args, world, rank, stream, pk, "bf16")
        if rank == 0:
            extra["inference"] = inf
    if rank == 0:
        if extra:
            line["configs"] = extra
        if world == 1 and not args.no_cpu_baseline:
            val, ms, cores, ph, kind = cpu_reference_run(5, 2)
            line["cpu_baseline"] = {"value": val, "unit": "triplets/s", "cores": cores, "kind": kind, "note": CPU_NOTE[kind],
                                    "sample": "B=%d items per step (1/%d of the GPU step), same K/N/C/Nn, 5 timed steps after 2 warm-up"
                                              % (CPU_SAMPLE_B, B // CPU_SAMPLE_B), "ms_per_step": ms, "phase_ms": ph}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="f16x3", choices=["tf32x3", "f16x3", "tf32", "bf16", "fp32_simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--materialised-gather", action="store_true", help="run K0 as its own kernel (materialised X operand) instead of gathering inside the GEMMs")
    ap.add_argument("--sampler", default="sharded", choices=["sharded", "global"],
                    help="sharded: every rank draws its own reference-exact stream(s) over its shard of the videos; "
                         "global: one stream of G*B items per step (SURVEY 8e), every rank takes its slice")
    ap.add_argument("--sampler-streams", type=int, default=0, help="reference-exact sampler streams per rank (0 = from the host core count)")
    ap.add_argument("--no-extra-configs", action="store_true", help="only the headline configuration (skip configs[2], [3], [4])")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

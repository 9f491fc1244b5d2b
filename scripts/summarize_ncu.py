"""Turn gpurun_out/*.ncu-rep + launch lists into the committed summaries under profiles/."""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def summarize(rep, out, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s\n# source: %s (ncu --set full --clock-control none --import-source on), one block per captured launch\n" % (tag, os.path.basename(rep)))
        for r in rows[2:]:
            f.write("\nkernel: %s\n" % r[hdr.index("Kernel Name")][:160])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("  %-72s %s %s\n" % (w, r[i], units[i]))


def traffic(rep):
    """{fc7_forward|wgrad: dram bytes read+write per launch} from a GEMM capture (forward = the fwd-epilogue template flag)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if "gemm_tc_kernel" not in name:
            continue
        args = name[name.index("Cfg<") + 4:].split(">")[0].replace("(bool)", "").replace("(int)", "").split(",")
        key = "fc7_forward" if args[6].strip() == "1" else "wgrad"
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * mult.get(units[i], 1.0)
        out[key] = tot
    return out


def launches(src, out, tag):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    with open(out, "w") as f:
        f.write("# %s\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n" % tag)
        f.write("# id, duration_us, kernel\n")
        tot = {}
        for r in rows[1:]:
            us = float(r[vi].replace(",", "")) / 1000.0
            name = r[ki].split("(")[0].replace("void ", "").replace("vv::<unnamed>::", "")[:90]
            f.write("%4d, %10.2f, %s\n" % (int(r[ii]), us, name))
            tot[name] = tot.get(name, 0) + us
        f.write("\n# share of all captured launches (steps + setup)\n")
        s = sum(tot.values())
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write("# %6.2f %%  %10.1f us  %s\n" % (100 * v / s, v, k))


if __name__ == "__main__":
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    go = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(os.path.join(go, rnd)):            # later rounds keep their captures in gpurun_out/<round>/
        go = os.path.join(go, rnd)
    pr = os.path.join(ROOT, "profiles")
    for p in ("f16x3", "tf32x3", "bf16"):
        if os.path.exists(os.path.join(go, "launches_%s.csv" % p)):
            launches(os.path.join(go, "launches_%s.csv" % p), os.path.join(pr, "%s_launches_%s.txt" % (rnd, p)),
                     "launch list, `python scripts/profile_step.py --precision %s --steps 4` (B=4096, K=4096, N=512, C=5, Nn=10)" % p)
    if os.path.exists(os.path.join(go, "launches_bench.csv")):
        launches(os.path.join(go, "launches_bench.csv"), os.path.join(pr, "%s_launches_bench.txt" % rnd),
                 "launch list of the bench command itself: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs` (value leg, per-kernel leg, e2e leg)")
    tr = {}
    for key, rep in (("f16x3_gathered", "prof_gemm_f16x3"), ("bf16_gathered", "prof_gemm_bf16")):
        p = os.path.join(go, rep + ".ncu-rep")
        if os.path.exists(p):
            tr[key] = traffic(p)
    if tr:
        import json
        json.dump(tr, open(os.path.join(pr, "traffic.json"), "w"), indent=1)
    for rep, tag in (("prof_gemm_f16x3", "fc7 forward + transposed wgrad GEMMs, f16x3, gather fused (default path)"),
                     ("prof_gemm_tf32x3", "fc7 forward + wgrad GEMMs, tf32x3"), ("prof_gemm_bf16", "fc7 forward + wgrad GEMMs, bf16, gather fused"),
                     ("prof_stream_f16x3", "streaming kernels (gather plan, fused rank loss, update), f16x3 step"),
                     ("prof_gather_f16x3", "K0 gather kernel (materialised path, --materialised), f16x3"),
                     ("prof_rank_ring", "rank_ring_kernel (VV_RANK_RING=2, 3 CTAs/SM), timed alone on cold data: scripts/bench_rank.py 2 3"),
                     ("prof_stream_tf32x3", "streaming kernels (gather, rank loss, update), tf32x3 step")):
        p = os.path.join(go, rep + ".ncu-rep")
        if os.path.exists(p):
            summarize(p, os.path.join(pr, "%s_%s.txt" % (rnd, rep)), tag)

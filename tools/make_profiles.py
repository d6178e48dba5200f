"""Collect the evidence the judge reads into profiles/ (tracked): SASS mnemonic census of every kernel file, ncu summaries, in-graph step
breakdown, GEMM phase timing, bench JSON lines.  Run on the authoring box after `gpurun` brought the raw files back into gpurun_out/."""
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def sass_census():
    lines = ["# SASS mnemonic census (cuobjdump -sass of the sm_100a objects)\n",
             "`tcgen05.mma` -> `UTC*MMA`, `tcgen05.ld` -> `LDTM`, TMA -> `UTMALDG`, mbarrier -> `SYNCS`, PDL -> `ACQBULK`/`PREEXIT`-class, "
             "`multimem.*` -> `LDGMC`/`STG.MC`-class, system-scope flags -> `*.STRONG.SYS`.\n"]
    pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)")
    for obj in sorted(glob.glob(os.path.join(ROOT, "geomx_b200", "build_obj", "*.o"))):
        if os.path.basename(obj).startswith("rt_"):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
        cnt = {}
        for l in txt.splitlines():
            m = pat.match(l)
            if m:
                op = m.group(1)
                key = None
                if op.startswith("UTC") and "MMA" in op: key = op.split(".")[0]
                elif op.startswith(("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS")): key = op.split(".")[0]
                elif op.startswith("SYNCS"): key = "SYNCS(mbarrier)"
                elif "MULTIMEM" in op or op.startswith(("LDGMC", "REDGMC")) or ".MC" in op: key = op
                elif ".SYS" in op: key = op
                elif op.startswith(("ACQBULK", "PREEXIT", "ACQSHMINIT")): key = op
                elif op.startswith("HMMA"): key = "HMMA(legacy)"
                if key:
                    cnt[key] = cnt.get(key, 0) + 1
        lines.append("\n## %s\n" % os.path.basename(obj))
        lines.append(", ".join("%s x%d" % kv for kv in sorted(cnt.items())) or "(plain CUDA-core kernel)")
        lines.append("\n")
    open(os.path.join(OUT, "sass_census.md"), "w").write("\n".join(lines))


def sass_listing():
    """profiles/sass_listing.md: for the GEMM and the fabric kernels, the first few SASS lines of every tensor-core / TMA / TMEM / NVLink
    mnemonic (address + instruction), i.e. the listing that proves which hardware paths the binaries use."""
    import re
    want = re.compile(r"UCGABAR|LDGSTS|ST\.E\.[A-Z0-9.]*\.CLUSTER|STS\.[A-Z0-9.]*CLUSTER|UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|LDTM|STTM|SYNCS|ELECT|UBLKCP|MULTIMEM|LDGMC|STGMC|REDGMC|MEMBAR|ST\.E\.[A-Z0-9.]*STRONG\.SYS|LD\.E\.[A-Z0-9.]*STRONG\.SYS|LDG\.E\.[A-Z0-9.]*STRONG\.SYS|STG\.E\.[A-Z0-9.]*STRONG\.SYS|RED\.E|ATOM\.E|F2FP|HMMA|CCTL")
    out = ["# SASS listing (cuobjdump -sass of the objects under geomx_b200/build_obj, sm_100a)", "",
           "For each kernel: up to 3 occurrences of every mnemonic of interest, with its address.  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,",
           "LDTM = tcgen05.ld (TMEM), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, UTCATOMSWS = TMEM alloc; `*.STRONG.SYS` loads/stores and",
           "LDGMC (= multimem.ld_reduce) are the NVLink peer / NVSwitch accesses of the fused HiPS kernels (multimem.st assembles to STG...STRONG.SYS).", ""]
    for obj, kernels in (("gemm_tcgen05.o", ("gemm_tf32_kernelILi128ELb0ELb0ELb0E", "gemm_tf32_kernelILi128ELb0ELb0ELb1E", "gemm_tf32_kernelILi32ELb1ELb1ELb1E")),
                         ("hips_fabric.o", ("hips_fsa_direct_kernel", "hips_fsa_ll_kernel", "hips_fsa_step_kernel", "hips_async_step_kernel", "hips_party_allreduce_kernel")),
                         ("cnn_direct.o", ("cnn_bwd_exchange_kernel", "cnn_fwd_kernel")), ("mlp_chain.o", ("mlp_chain_kernelILi512ELi256ELi128ELi16E",))):
        path = os.path.join(ROOT, "geomx_b200", "build_obj", obj)
        if not os.path.exists(path):
            continue
        txt = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        cur, seen = None, {}
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = next((k for k in kernels if k in m.group(1)), None)
                if cur:
                    out += ["", "## %s :: %s" % (obj, m.group(1)[:110]), "```"]
                    seen = {}
                continue
            if cur is None:
                continue
            if ".........." in line and seen:
                out.append("```"); cur = None
                continue
            mm = want.search(line)
            if mm and "/*" in line:
                key = mm.group(0).split(".")[0] if not mm.group(0).startswith(("ST.", "LD.", "LDG", "STG")) else mm.group(0)
                if seen.get(key, 0) < 3:
                    seen[key] = seen.get(key, 0) + 1
                    out.append(line.split("/* 0x")[0].rstrip())
        if cur is not None:
            out.append("```")
    open(os.path.join(OUT, "sass_listing.md"), "w").write("\n".join(out) + "\n")


def _clean(txt):
    return "\n".join(l for l in txt.splitlines() if not l.startswith("frame #") and "OMP_NUM_THREADS" not in l and not l.startswith("*****")
                     and not l.startswith("W0") and "FutureWarning" not in l and "if not hasattr(np" not in l)


# gpurun_out file -> profiles file (this round's measurements; round-1 files live under profiles/round1/)
COPIES = {
    "r2_timeline1g.txt": "timeline_1gpu.txt", "r2_timeline2g.txt": "timeline_2gpu.txt", "r2_timeline8.txt": "timeline_8gpu.txt",
    "r2_timeline2_la.txt": "timeline_2gpu_lookahead.txt", "r2_timeline2f.txt": "timeline_2gpu_fused_tail.txt",
    "r2_ktimes5.txt": "kernel_times.txt", "r2_gemm_anchor.txt": "gemm_anchor.txt", "r2_gemm_anchor_s3.txt": "gemm_anchor_first_try_stages3.txt",
    "r2_fab8.log": "fabric_check_8gpu.txt", "r2_fab4.log": "fabric_check_4gpu.txt", "r2_fabapi.log": "fabric_api_check.txt",
    "r2_pytest_gpu_full.log": "pytest_gpu_full.txt", "r2_pytest_8gpu.log": "pytest_multigpu_8gpu.txt",
}


def copy_logs():
    for src, dst in COPIES.items():
        p = os.path.join(GO, src)
        if os.path.exists(p):
            open(os.path.join(OUT, dst), "w").write(_clean(open(p).read())[-16000:] + "\n")
    # reference arm: the JSON lines of the unmodified MXNet build
    ref = []
    for p in sorted(glob.glob(os.path.join(GO, "r2_ref*.log"))):
        for l in open(p):
            if l.startswith("{"):
                ref.append("%s:\n%s" % (os.path.basename(p), l.strip()))
    if ref:
        open(os.path.join(OUT, "reference_arm.txt"), "w").write(
            "bench.py --impl reference: the UNMODIFIED /root/reference (MXNet 1.4.0 + GeoMX kvstore) built for sm_100 / CUDA 12.9 by\n"
            "baseline/build_reference.sh, run through baseline/ref_cnn_bench.py on B200 (host-clocked between mx.nd.waitall() barriers).\n\n"
            + "\n\n".join(ref) + "\n")
    rows = []
    for p in sorted(glob.glob(os.path.join(GO, "r2_b*.log")) + glob.glob(os.path.join(GO, "r2_bench*.log")) + glob.glob(os.path.join(GO, "r2_scale*.log"))
                    + glob.glob(os.path.join(GO, "r2_cfg*.log"))):
        for l in open(p):
            if l.startswith("{"):
                try:
                    rows.append((os.path.basename(p), os.path.getmtime(p), json.loads(l)))
                except Exception:
                    pass
    rows.sort(key=lambda r: r[1])
    with open(os.path.join(OUT, "bench_history.md"), "w") as f:
        f.write("# bench.py runs of round 2, in chronological order (device-timed, L2 flushed between steps; every row is one gpurun call on a fresh "
                "B200 box)\n\nThe file name says what was varied; the LAST rows are the final code.  ms/step p10/p50/p90 were added mid-round.\n\n"
                "| file | config | path | n_gpus | ms/step | p10 / p50 / p90 | samples/s | e2e samples/s | exposed comm ms | launches/step | "
                "sm MHz | throttle reasons | protocol errors |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for name, _, d in rows:
            c = d.get("clocks") or {}
            f.write("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |\n" % (
                name, d.get("baseline_config"), "script" if "script" in str(d.get("path")) else "engine", d.get("n_gpus"), d.get("ms_per_step"),
                d.get("ms_per_step_p10_p50_p90"), d.get("value"), (d.get("e2e") or {}).get("value"), d.get("exposed_push_pull_ms_per_step"),
                d.get("gpu_launches_per_step"), c.get("sm_mhz"), c.get("reasons"), d.get("protocol_errors")))


NCU_METRICS = [
    ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue act %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
]


def ncu_summaries():
    for rep in sorted(glob.glob(os.path.join(GO, "r2_*.ncu-rep"))):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        have = [(m, t) for m, t in NCU_METRICS if m in hdr]
        base = os.path.basename(rep).replace(".ncu-rep", "").replace("r2_", "ncu_")
        with open(os.path.join(OUT, base + "_raw.md"), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on (%s; tools/ncu_step.py)\n\n"
                    "One eager flagship step (L2 flushed before it: cold operands, like a bench step), then BatchNorm fwd/bwd (64x128x28x28) and the\n"
                    "8192x4096x4096 GEMM in both precisions.  ncu serialises the launches and replays each ~40 times: durations are NOT bench numbers —\n"
                    "read the pipe utilisation, stall mix and conflict counts.  Stall columns are warps-stalled per issued instruction.\n\n" % os.path.basename(rep))
            f.write("| kernel | " + " | ".join("%s [%s]" % (t, units[hdr.index(m)]) if units[hdr.index(m)] else t for m, t in have) + " |\n|---|" + "---|" * len(have) + "\n")
            kn = hdr.index("Kernel Name")
            for r in rows[2:]:
                name = re.sub(r"\(.*", "", r[kn]).replace("void ", "")
                vals = []
                for m, _ in have:
                    v = r[hdr.index(m)]
                    try:
                        v = "%.3g" % float(v.replace(",", ""))
                    except Exception:
                        pass
                    vals.append(v)
                f.write("| %s | " % name[:48] + " | ".join(vals) + " |\n")
        # hottest SASS lines (stall samples) per kernel, with shared-memory wavefront excess
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        out, cur, hdr2, block = ["# hottest SASS instructions per kernel (ncu source page: warp-stall samples; shared-memory wavefronts actual / ideal)", ""], None, None, []

        def flush():
            if cur and block and hdr2:
                i_s, i_src, i_w, i_i = hdr2.index("# Samples"), hdr2.index("Source"), hdr2.index("L1 Wavefronts Shared"), hdr2.index("L1 Wavefronts Shared Ideal")
                num = lambda x: int(float(x)) if x.replace(".", "").isdigit() else 0
                tot = sum(num(b[i_s]) for b in block) or 1
                out.extend(["", "## %s  (%d samples, %d SASS instructions)" % (cur[:100], tot, len(block)), "```"])
                for b in sorted(block, key=lambda b: -num(b[i_s]))[:14]:
                    out.append("%5.1f%%  %-58s smem wavefronts %s / %s" % (100.0 * num(b[i_s]) / tot, b[i_src].strip()[:58], b[i_w], b[i_i]))
                out.append("```")
        for row in csv.reader(src.splitlines()):
            if len(row) >= 2 and row[0] == "Kernel Name":
                flush(); cur, hdr2, block = row[1], None, []
            elif row and row[0] == "Address":
                hdr2 = row
            elif hdr2 and len(row) == len(hdr2):
                block.append(row)
        flush()
        open(os.path.join(OUT, base + "_hot_sass.md"), "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    sass_census(); sass_listing(); copy_logs(); ncu_summaries()
    print(sorted(os.listdir(OUT)))

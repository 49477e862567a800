"""Algorithmic bytes per tcgen05 GEMM / implicit-conv launch of one forward, from the dispatcher log (GGML_B200_GEMM_LOG=1 ... | sort | uniq -c):
sum over the launches of ((M + N) * K * 2 + M * N * 4) * batch (16-bit operands read once, f32 result written once) / launch count.
usage: gemm_algorithmic_bytes.py <uniq -c'd GEMMLOG> <metrics.json to annotate>"""
import json
import re
import sys

tot = n = 0
for line in open(sys.argv[1]):
    m = re.match(r"\s*(\d+) GEMMLOG \S+ \S+ M (\d+) N (\d+) K (\d+) batch (\d+)", line)
    if not m:
        continue
    c, M, N, K, b = map(int, m.groups())
    tot += c * ((M + N) * K * 2 + M * N * 4) * b
    n += c
d = json.load(open(sys.argv[2]))
d["gemm_algorithmic_bytes_per_launch"] = tot / n
d["gemm_algorithmic_note"] = (f"sum over the {n} GEMM / implicit-conv launches of one batched forward of ((M + N) * K * 2 + M * N * 4) * batch bytes (16-bit operands "
                              "read once, f32 result written once), divided by the launch count; shapes from the dispatcher log (GGML_B200_GEMM_LOG)")
json.dump(d, open(sys.argv[2], "w"), indent=1)
print(n, tot / n)

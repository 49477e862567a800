"""Turn an `ncu --csv --metrics ...` log over ALL launches of one model call into a per-kernel table (JSON + markdown).

usage: ncu_kernel_table.py <ncu.csv> <out.json> <out.md> "<title>" [first_kernel_regex]
  ncu command that produces the input (one GPU, never under a timed run):
    GGML_B200_CUDA_GRAPHS=0 ncu --clock-control none --csv --log-file x.csv --metrics \
      gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct  python scripts/one_forward.py <case> 2
The LAST complete model call in the log is summarised (delimited by `first_kernel_regex`, default: the k_timestep_embedding launch that
starts every UNet / DiT forward; for the VAE pass 'k_im2col' -- its conv_in is the only materialised im2col and comes first).
Per kernel: launches, total / average duration, DRAM bytes (read + write), achieved DRAM GB/s over the kernel's own duration, tensor-pipe
and warp occupancy averages (duration-weighted), L2 hit rate.  ncu serialises launches and runs them cold-cache: compare SHARES and
per-launch bytes, not absolute step time.
"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "%": 1.0}


def num(v):
    return float(str(v).replace(",", ""))


def main():
    path, out_json, out_md, title = sys.argv[1:5]
    first = re.compile(sys.argv[5] if len(sys.argv) > 5 else "timestep")
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        L = launches.setdefault(int(r["ID"]), dict(name=r["Kernel Name"], grid=r["Grid Size"], block=r["Block Size"]))
        L[r["Metric Name"]] = num(r["Metric Value"]) * UNIT.get(r["Metric Unit"], 1.0)
    rows = list(launches.values())
    idx = [i for i, r in enumerate(rows) if first.search(r["name"])]
    bounds = idx + [len(rows)]
    segs = [rows[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)]
    longest = max((len(s) for s in segs), default=0)
    full = [s for s in segs if len(s) * 2 >= longest]
    # NCU_SEGMENT=-2: the segment BEFORE the last -- with projections hoisted to the start of a graph (side streams) the launches of call
    # k + 1 that precede its timestep embedding land at the end of segment k, so only a segment followed by another call is complete
    import os
    which = int(os.environ.get("NCU_SEGMENT", "-1"))
    seg = (full[which] if len(full) >= abs(which) else full[-1]) if full else rows
    agg = collections.OrderedDict()
    for r in seg:
        n = re.sub(r"\(.*", "", r["name"])
        n = re.sub(r"void |\(anonymous namespace\)::|<unnamed>::", "", n)
        a = agg.setdefault(n, dict(kernel=n, launches=0, us=0.0, dram_bytes=0.0, dram_read=0.0, dram_write=0.0, tensor_w=0.0, warps_w=0.0, l2hit_w=0.0))
        t = r.get("gpu__time_duration.sum", 0.0)
        a["launches"] += 1
        a["us"] += t
        a["dram_read"] += r.get("dram__bytes_read.sum", 0.0)
        a["dram_write"] += r.get("dram__bytes_write.sum", 0.0)
        a["tensor_w"] += t * r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a["warps_w"] += t * r.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
        a["l2hit_w"] += t * r.get("lts__t_sector_hit_rate.pct", 0.0)
    tot = sum(a["us"] for a in agg.values())
    kernels = []
    for a in sorted(agg.values(), key=lambda a: -a["us"]):
        a["dram_bytes"] = a["dram_read"] + a["dram_write"]
        kernels.append(dict(kernel=a["kernel"], launches=a["launches"], total_us=a["us"], avg_us=a["us"] / a["launches"], share=a["us"] / tot,
                            dram_bytes=a["dram_bytes"], dram_read_bytes=a["dram_read"], dram_write_bytes=a["dram_write"],
                            dram_bytes_per_launch=a["dram_bytes"] / a["launches"], dram_gbs=a["dram_bytes"] / max(a["us"], 1e-9) / 1e3,
                            tensor_pipe_pct=a["tensor_w"] / max(a["us"], 1e-9), warps_active_pct=a["warps_w"] / max(a["us"], 1e-9),
                            l2_hit_pct=a["l2hit_w"] / max(a["us"], 1e-9)))
    out = dict(title=title, source=path, launches=len(seg), sum_us=tot, dram_bytes_total=sum(k["dram_bytes"] for k in kernels), kernels=kernels,
               note="ncu metrics pass: launches serialised, cold cache, default clocks (--clock-control none); dram_gbs = dram bytes / kernel duration")
    json.dump(out, open(out_json, "w"), indent=1)
    md = [f"# {title}", f"source: `{path}` (ncu metrics pass over every launch of one call; serialised, cold-cache: compare shares and bytes)",
          f"launches {len(seg)}, sum of kernel durations {tot / 1e3:.2f} ms, DRAM traffic {out['dram_bytes_total'] / 1e9:.2f} GB", "",
          "| kernel | launches | total us | avg us | share | DRAM MB/launch | DRAM GB/s | tensor pipe % | warps active % | L2 hit % |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for k in kernels:
        md.append(f"| `{k['kernel'][:70]}` | {k['launches']} | {k['total_us']:.1f} | {k['avg_us']:.1f} | {100 * k['share']:.1f}% | {k['dram_bytes_per_launch'] / 1e6:.2f} | "
                  f"{k['dram_gbs']:.0f} | {k['tensor_pipe_pct']:.1f} | {k['warps_active_pct']:.1f} | {k['l2_hit_pct']:.1f} |")
    open(out_md, "w").write("\n".join(md) + "\n")
    print("\n".join(md[:30]))


if __name__ == "__main__":
    main()

"""Summarise `ncu --set full` reports (.ncu-rep, read HERE with `ncu -i`) into one committed JSON: per captured launch the numbers the
roofline argument needs -- duration, DRAM bytes and achieved GB/s, tensor-pipe %, occupancy, L2 hit rate, registers, shared memory -- and
the five most-sampled source lines.   usage: ncu_full_summary.py out.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__shared_mem_config_size", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = (r[i] + (" " + units[i] if units[i] else "")).strip()
        def val(k):
            if k not in hdr:
                return 0.0
            i = hdr.index(k)
            try:
                return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
            except ValueError:
                return 0.0
        us = val("gpu__time_duration.sum")
        b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        d["achieved_dram_gbs"] = round(b / max(us, 1e-9) / 1e3, 1)
        d["l2_to_sm_gbs"] = round(val("l1tex__m_xbar2l1tex_read_bytes.sum") / max(us, 1e-9) / 1e3, 1)
        res.append(d)
    return res


def top_lines(rep, launch):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{launch}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr = rows[1]
    if "Source" not in hdr or "# Samples" not in hdr:
        return []
    iS, iN = hdr.index("Source"), hdr.index("# Samples")
    seen, tot = {}, 0
    for r in rows[2:]:
        if len(r) <= iN:
            continue
        try:
            n = int(r[iN])
        except ValueError:
            continue
        key = (r[0], r[iS])
        if key in seen:
            continue
        seen[key] = n
        tot += n
    top = sorted(seen.items(), key=lambda kv: -kv[1])[:5]
    return [dict(sass=k[1][:90], samples_pct=round(100.0 * n / max(tot, 1), 1)) for k, n in top]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    res = []
    for rep in reps:
        for i, d in enumerate(raw(rep)):
            d["source"] = rep
            d["top_sampled_sass"] = top_lines(rep, i + 1)
            res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["Kernel Name"][:60], d.get("gpu__time_duration.sum"), "dram GB/s", d["achieved_dram_gbs"], "tensor %",
              d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main()

"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum and optionally dram bytes) per
kernel: launches, total device time, share, DRAM traffic.  Writes a markdown table and, with
--traffic-json, the per-kernel DRAM bytes bench.py reports as `roofline.traffic`."""
import argparse
import collections
import csv
import json
import re

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--md")
ap.add_argument("--traffic-json")
ap.add_argument("--build-id", default="")
a = ap.parse_args()

rows = list(csv.reader(open(a.csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
NAMES = {"gemm_sm100_kernel<0>": "gemm_std", "gemm_sm100_kernel<(int)0>": "gemm_std",
         "gemm_sm100_kernel<1>": "gemm_l2norm", "gemm_sm100_kernel<(int)1>": "gemm_l2norm",
         "gemm_sm100_kernel<2>": "gemm_stft", "gemm_sm100_kernel<(int)2>": "gemm_stft",
         "gemm_sm100_kernel<3>": "gemm_head", "gemm_sm100_kernel<(int)3>": "gemm_head"}


EPI = {0: "gemm_std", 1: "gemm_l2norm", 2: "gemm_stft", 3: "gemm_head", 4: "gemm_std_precise", 5: "gemm_l2norm_precise",
       6: "gemm_stft_precise"}


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("wv::", "").strip()
    m = re.search(r"gemm_sm100_kernel<([^>]*)>", name)
    if m:   # template arguments: epilogue id, CTA-pair flag ("0, 1", "(int)0, (bool)1", "0, true" ...)
        args = [re.sub(r"\([a-z ]+\)", "", t).strip() for t in m.group(1).split(",")]
        epi = EPI.get(int(args[0]), m.group(0))
        pair = len(args) > 1 and args[1] in ("1", "true")
        return epi + ("_cta_pair" if pair else "")
    return n.replace("_kernel", "")


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)


agg = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
for r in data:
    if len(r) <= vi:
        continue
    k = short(r[ki])
    if r[mi] == "gpu__time_duration.sum":
        agg[k]["launches"] += 1
        agg[k]["us"] += to_us(r[vi], r[ui])
    elif r[mi] == "dram__bytes_read.sum":
        agg[k]["rd"] += to_bytes(r[vi], r[ui])
    elif r[mi] == "dram__bytes_write.sum":
        agg[k]["wr"] += to_bytes(r[vi], r[ui])
tot = sum(v["us"] for v in agg.values()) or 1.0
lines = ["| kernel | launches | device time (us) | share | DRAM read (MB) | DRAM write (MB) |", "|---|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    lines.append(f"| {k} | {v['launches']} | {v['us']:.1f} | {100 * v['us'] / tot:.1f} % | {v['rd'] / 1e6:.1f} | {v['wr'] / 1e6:.1f} |")
lines.append(f"| **total** | {sum(v['launches'] for v in agg.values())} | {tot:.1f} | 100 % | | |")
out = "\n".join(lines)
print(out)
if a.md:
    open(a.md, "w").write(out + "\n")
if a.traffic_json:
    # bench.py's kernel classes do not distinguish the CTA-pair instantiation: fold it into its class
    tj = collections.defaultdict(lambda: {"dram_bytes_per_step": 0.0, "launches": 0})
    for k, v in agg.items():
        if v["rd"] + v["wr"] > 0:
            e = tj[k.replace("_cta_pair", "")]
            e["dram_bytes_per_step"] += v["rd"] + v["wr"]; e["launches"] += v["launches"]
    out_j = {k: {**v, "dram_bytes_per_launch": v["dram_bytes_per_step"] / max(1, v["launches"])} for k, v in tj.items()}
    out_j["_dram_bytes_per_step"] = sum(v["rd"] + v["wr"] for v in agg.values())
    out_j["_device_time_us_per_step_ncu"] = tot
    if a.build_id:
        out_j["_build_id"] = a.build_id
    json.dump(out_j, open(a.traffic_json, "w"), indent=1)

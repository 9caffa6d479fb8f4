"""ncu launch lists of probes/unet_conv_launches.py -> the build-stamped traffic record bench.py reads.

    python probes/make_traffic_json.py <train.csv> [<predict.csv>] > profiles/r02_traffic.json

Only the tcgen05 convolution kernels (conv_umma_*) are summed; the weight-pack / channel-pad helpers and the L2-flush
memsets that the probe launches around them are listed separately."""
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_stamp


def parse(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    idx = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        name, metric, unit = r[idx["Kernel Name"]], r[idx["Metric Name"]], r[idx["Metric Unit"]]
        val = float(r[idx["Metric Value"]].replace(",", ""))
        if metric == "gpu__time_duration.sum":
            val *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0}.get(unit, 1e-6)
        else:
            val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d = per.setdefault(r[idx["ID"]], {"name": name.split("(")[0].replace("void ", "").replace("b200::", ""), "ms": 0, "rd": 0, "wr": 0})
        d[{"gpu__time_duration.sum": "ms", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr"}[metric]] += val
    return list(per.values())


def family(launches):
    conv = [l for l in launches if l["name"].startswith(("conv_umma", "conv_ws"))]
    other = [l for l in launches if not l["name"].startswith(("conv_umma", "conv_ws"))]
    return {"launches": len(conv), "dram_bytes_per_step": sum(l["rd"] + l["wr"] for l in conv),
            "dram_read_bytes": sum(l["rd"] for l in conv), "dram_write_bytes": sum(l["wr"] for l in conv),
            "ncu_ms_cold_serialised": round(sum(l["ms"] for l in conv), 4),
            "by_kernel": {k: {"launches": sum(1 for l in conv if l["name"] == k),
                              "dram_bytes": sum(l["rd"] + l["wr"] for l in conv if l["name"] == k),
                              "ms": round(sum(l["ms"] for l in conv if l["name"] == k), 4)} for k in sorted({l["name"] for l in conv})},
            "helper_launches_excluded": sorted({l["name"] for l in other})}


out = {"kernel_stamp": kernel_stamp(), "source": "profiles/r02_traffic.json <- ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
       "dram__bytes_write.sum --clock-control none on probes/unet_conv_launches.py (L2 flushed before every launch)", "families": {}}
out["families"]["conv_fprop_dgrad"] = family(parse(sys.argv[1]))
if len(sys.argv) > 2:
    fam = family(parse(sys.argv[2]))
    scale = 147 / 16.0      # one batch of 16 patches was profiled; a 512x512x256 volume is 147 patches
    fam["dram_bytes_per_batch16"] = fam["dram_bytes_per_step"]
    fam["dram_bytes_per_step"] = fam["dram_bytes_per_step"] * scale
    out["families"]["predict_conv_fprop"] = fam
print(json.dumps(out, indent=1))

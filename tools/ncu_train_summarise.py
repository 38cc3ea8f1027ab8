"""ncu --csv --log-file output of the training step (tools/gpu_evidence.sh step 4) -> per-kernel summary JSON:
python tools/ncu_train_summarise.py gpurun_out/r02_ncu_train_kernels.csv profiles/r02_ncu_train_kernels.json"""
import csv
import json
import sys
from collections import defaultdict


def main():
    src, out = sys.argv[1:3]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    c = {h: i for i, h in enumerate(hdr)}
    per = defaultdict(lambda: defaultdict(dict))
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        key = (r[c["ID"]], r[c["Kernel Name"]].split("(")[0])
        try:
            per[key][r[c["Metric Name"]]] = (float(r[c["Metric Value"]].replace(",", "")), r[c["Metric Unit"]])
        except ValueError:
            pass
    agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_MB": 0.0, "dram_pct_weighted": 0.0})
    scale_t = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    scale_b = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    for (_, name), m in per.items():
        t = m.get("gpu__time_duration.sum")
        if not t:
            continue
        us = t[0] * scale_t.get(t[1], 1.0)
        a = agg[name]
        a["launches"] += 1
        a["us"] += us
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in m:
                a["dram_MB"] += m[k][0] * scale_b.get(m[k][1], 1e-6)
        if "dram__throughput.avg.pct_of_peak_sustained_elapsed" in m:
            a["dram_pct_weighted"] += m["dram__throughput.avg.pct_of_peak_sustained_elapsed"][0] * us
    res = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        res.append({"kernel": name, "launches": a["launches"], "total_us": round(a["us"], 1),
                    "dram_GB": round(a["dram_MB"] / 1e3, 3),
                    "achieved_GBps": round(a["dram_MB"] / 1e3 / (a["us"] * 1e-6), 1) if a["us"] else None,
                    "dram_pct_of_peak_time_weighted": round(a["dram_pct_weighted"] / a["us"], 1) if a["us"] else None})
    json.dump({"what": "one native training step (batch 2 x 128^3) under ncu, loss / BatchNorm / optimizer / weight-gradient "
                       "kernels only; DRAM bytes and throughput from ncu", "kernels": res}, open(out, "w"), indent=1)
    for r in res[:12]:
        print(r)


if __name__ == "__main__":
    main()

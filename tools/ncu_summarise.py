"""ncu raw CSV (one row per launch, plan order) + step list -> compact per-launch JSON summary."""
import csv
import json
import sys


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    raw, steps_json, out = sys.argv[1:4]
    rows = list(csv.reader(open(raw)))
    hdr, data = rows[0], rows[2:]
    steps = json.load(open(steps_json))
    col = {h: i for i, h in enumerate(hdr)}
    want = {"dur_us": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
            "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "issue_pct": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
            "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
            "smem_dyn_kb": "launch__shared_mem_per_block_dynamic"}
    units = rows[1]
    res = []
    for r, st in zip(data, steps["steps"]):
        e = {"step": st["name"], "kernel": r[col["Kernel Name"]].split("(")[0][:40], "alg_MB": round(st["bytes"] / 1e6, 2),
             "alg_GFLOP": round(st["flops"] / 1e9, 3)}
        for k, m in want.items():
            if m in col:
                v = num(r[col[m]])
                u = units[col[m]]
                if v is not None and k.startswith("dram_r") or k.startswith("dram_w"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) if v is not None else None
                    k = k + "_MB"
                if v is not None and k == "dur_us":
                    v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
                e[k] = round(v, 3) if v is not None else None
        res.append(e)
    json.dump({"group": steps["group"], "launches": res}, open(out, "w"), indent=0)
    tot = sum(e.get("dur_us") or 0 for e in res)
    print(f"{len(res)} launches, {tot:.1f} us under ncu ({tot / steps['group']:.1f} us per window)")


if __name__ == "__main__":
    main()

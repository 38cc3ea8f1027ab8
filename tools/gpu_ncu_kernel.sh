#!/bin/bash
# ncu --set full of ONE launch of a kernel (regex $1) of the single-window plan; key metrics -> gpurun_out/$2_*.txt
K=$1; T=${2:-k}; O=gpurun_out; mkdir -p $O
PROFILE_GROUP=${PROFILE_GROUP:-1} timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$K --launch-count 1 -f -o /tmp/$T python tools/ncu_plan.py /tmp/steps_$T.json > $O/${T}_ncu.log 2>&1
ncu -i /tmp/$T.ncu-rep --page details > $O/${T}_details.txt 2>/dev/null
ncu -i /tmp/$T.ncu-rep --page raw --csv > $O/${T}_raw.csv 2>/dev/null
ncu -i /tmp/$T.ncu-rep --page source --csv > /tmp/${T}_source.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("/tmp/${T}_source.csv")))
if rows:
    hdr = rows[0]
    si = [i for i, h in enumerate(hdr) if "Sampling" in h or "Samples" in h]
    if si:
        s0 = si[0]
        body = [r for r in rows[1:] if len(r) > s0 and r[s0].replace(",", "").isdigit()]
        body.sort(key=lambda r: -int(r[s0].replace(",", "")))
        w = csv.writer(open("$O/${T}_source_hot.csv", "w")); w.writerow(hdr)
        for r in body[:50]: w.writerow(r)
PY
grep -E "Duration|Elapsed Cycles|Registers Per|Theoretical Occ|Achieved Occ|Executed Ipc|Issue Slots Busy|No Eligible|Eligible Warps|DRAM Throughput|Memory Throughput|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth|Shared Memory Con|Local" $O/${T}_details.txt | head -40

#!/bin/bash
# One B200: tools/gpu_validate.sh, then a finer (XT, YT, ring depth) sweep of the heavy layers
bash tools/gpu_validate.sh c
AUTOTUNE_ONLY=dec1.unit0,enc1.unit1,dec1.att.conv1,dec0.att.conv1,enc0.unit1,enc1.unit0,dec2.unit0,dec2.att.conv1,enc2.unit,dec3.unit0,dec3.att.conv1,up1,up0,down0 AUTOTUNE_XT=1,2,3,4,6,8,16 AUTOTUNE_YT=1,2,3,4,6,8,16 AUTOTUNE_NST=0,2,3 PROFILE_GROUP=8 timeout 400 python tools/autotune_tiles.py gpurun_out/c_autotune_fine.tsv 2> gpurun_out/c_autotune_fine.err | cut -c1-300

O=gpurun_out; mkdir -p $O
PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/yt_default.tsv > /dev/null 2>&1
VSSEG_TC_YT_MAX=2 PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/yt_max2.tsv > /dev/null 2>&1
VSSEG_TC_YT_MAX=1 PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/yt_max1.tsv > /dev/null 2>&1
VSSEG_TC_FILL_LAT=1500 PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/yt_lat1500.tsv > /dev/null 2>&1
VSSEG_TC_FILL_BPC=24 PROFILE_GROUP=8 timeout 300 python tools/profile_plan.py $O/yt_bpc24.tsv > /dev/null 2>&1
tail -1 $O/yt_*.tsv

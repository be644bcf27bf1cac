O=gpurun_out/r2ab; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_c5.csv python tools/dbg_map.py - C5 4000000 > $O/dbg_c5.txt 2>&1
ls $O

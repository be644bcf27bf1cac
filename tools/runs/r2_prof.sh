# ncu --set full of one kernel (regex $1) of the C4 1M-pair debug run with options $2
O=gpurun_out/r2p; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-2} -c ${4:-1} -f -o $O/prof_$5 python tools/dbg_map.py "$2" ${6:-C4} 1000000 > $O/ncu_$5.log 2>&1
ls -la $O

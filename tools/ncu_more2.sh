P=r2x
mkdir -p gpurun_out
cap() { python tools/ncu_case.py $2 4 $4 > gpurun_out/${P}_$1_case.json 2> gpurun_out/${P}_$1_case.err && ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -o gpurun_out/${P}_$1 python tools/ncu_case.py $2 4 $4 > gpurun_out/${P}_$1_ncu.log 2>&1; cat gpurun_out/${P}_$1_case.json; }
cap k2_lean2_fixed 1080p_dense_512_fixed dda_accumulate_lean2
cap k2_lean4_scatter60 1080p_scatter60_512 dda_accumulate_lean4

for d in 0 4; do
ERTGPU_R900_CHAIN=pipe ERTGPU_R900_DBG=$d ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__cycles_active.avg --clock-control none -k regex:r900_chain2 -s 4 -c 1 --csv --log-file gpurun_out/j_dbg$d.csv python tools/quickbench.py r900 72 4 > gpurun_out/j_qb.log 2>&1
done

# ncu counters that were still missing (round-1 verdict: "ncu traffic for cfg 3 / cfg 5"): one step of the 256-robot bound
# group with resident inputs (what bench.py's `value` times: four batched launches for all 256 robots) and the boxed-in
# variant of config 3.  Run under gpurun (one GPU);
# afterwards: cp gpurun_out/counters/*.csv profiles/r02c_counters/ && python scripts/counters_to_json.py r02c_counters
O=gpurun_out/counters; mkdir -p $O
M=smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size,launch__block_size
python scripts/run_robots.py --cycles 12 > $O/plain_robots.log 2>&1 && \
ncu --metrics $M --clock-control none -k regex:_batch_kernel --launch-skip 40 -c 4 --csv --log-file $O/robots_256_2000x56.csv python scripts/run_robots.py --cycles 12 > $O/robots_256_2000x56.log 2>&1
python scripts/run_workload.py --workload obstacles_dense_16384x56 --cycles 32 --resident > $O/plain_dense.log 2>&1 && \
ncu --metrics $M --clock-control none --launch-skip 90 -c 3 --csv --log-file $O/obstacles_dense_16384x56.csv python scripts/run_workload.py --workload obstacles_dense_16384x56 --cycles 32 --resident > $O/obstacles_dense_16384x56.log 2>&1
tail -n 2 $O/*.log

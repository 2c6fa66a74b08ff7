# round-2 (last session), fifth pass: warp-compacted footprint checks in the stream rollout.  Timing of the three stream
# workloads, full GPU test suite, ncu counters of the boxed-in config 3, the bench line.
# bash scripts/collect_evidence_r02c5.sh   (under gpurun, one GPU; outputs in gpurun_out/r02c/)
O=gpurun_out/r02c; mkdir -p $O gpurun_out/counters
(python scripts/time_variants.py --workload obstacles_dense_16384x56 --cycles 20; python scripts/time_variants.py --workload obstacles_16384x56 --cycles 20; python scripts/time_variants.py --flush --cycles 16; python scripts/run_robots.py --cycles 12) > $O/compaction_timing_final.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/test_gpu_1gpu.log
M=smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size,launch__block_size
ncu --metrics $M --clock-control none --launch-skip 90 -c 3 --csv --log-file gpurun_out/counters/obstacles_dense_16384x56.csv python scripts/run_workload.py --workload obstacles_dense_16384x56 --cycles 32 --resident > gpurun_out/counters/obstacles_dense_16384x56.log 2>&1
cp gpurun_out/counters/obstacles_dense_16384x56.csv profiles/r02c_counters/
python scripts/counters_to_json.py r02c_counters > $O/counters_json.log 2>&1
cp profiles/ncu_kernel_counters.json $O/ncu_kernel_counters.json
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv -lms 100 > $O/clocks_during_bench_default.csv 2>/dev/null &
SMI=$!
timeout 600 python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default_1gpu.err
kill $SMI
cat $O/test_gpu_1gpu.log $O/compaction_timing_final.txt; tail -n 3 $O/bench_default_1gpu.err

# round-2 (last session), fourth pass: closed-form line cells in the footprint check.  Parity tests that reach the footprint
# code, the boxed-in config 3 (warm, resident), its ncu counters, the full suite, the bench line.
# bash scripts/collect_evidence_r02c4.sh   (under gpurun, one GPU; outputs in gpurun_out/r02c/)
O=gpurun_out/r02c; mkdir -p $O
timeout 120 python scripts/run_workload.py --workload obstacles_dense_16384x56 --cycles 16 --resident > $O/dense_resident_warm.log 2>&1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/test_gpu_1gpu.log
M=smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size,launch__block_size
mkdir -p gpurun_out/counters
ncu --metrics $M --clock-control none --launch-skip 90 -c 3 --csv --log-file gpurun_out/counters/obstacles_dense_16384x56.csv python scripts/run_workload.py --workload obstacles_dense_16384x56 --cycles 32 --resident > gpurun_out/counters/obstacles_dense_16384x56.log 2>&1
cp gpurun_out/counters/obstacles_dense_16384x56.csv profiles/r02c_counters/
python scripts/counters_to_json.py r02c_counters > $O/counters_json.log 2>&1
cp profiles/ncu_kernel_counters.json $O/ncu_kernel_counters.json
timeout 600 python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default_1gpu.err
cat $O/test_gpu_1gpu.log $O/dense_resident_warm.log; tail -n 3 $O/bench_default_1gpu.err

# round-2 (last session), third pass: footprint check with batched line loads.  Full GPU test suite, the boxed-in config 3
# before/after (warm, resident), the missing ncu counters, then the bench line that reads them.
# bash scripts/collect_evidence_r02c3.sh   (under gpurun, one GPU; outputs in gpurun_out/r02c/)
O=gpurun_out/r02c; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/test_gpu_1gpu.log
timeout 120 python scripts/run_workload.py --workload obstacles_dense_16384x56 --cycles 16 --resident > $O/dense_resident_warm.log 2>&1
timeout 400 bash scripts/collect_counters_r02c.sh > $O/counters.log 2>&1
mkdir -p profiles/r02c_counters && cp gpurun_out/counters/robots_256_2000x56.csv gpurun_out/counters/obstacles_dense_16384x56.csv profiles/r02c_counters/ 2>/dev/null
python scripts/counters_to_json.py r02c_counters > $O/counters_json.log 2>&1
cp profiles/ncu_kernel_counters.json $O/ncu_kernel_counters.json
timeout 600 python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default_1gpu.err
cat $O/test_gpu_1gpu.log $O/dense_resident_warm.log; tail -n 3 $O/counters.log $O/bench_default_1gpu.err

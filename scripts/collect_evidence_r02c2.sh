# round-2 (last session), second pass: new parity tests, the missing ncu counters (robots step, boxed-in config 3), then the
# bench line that reads them.  bash scripts/collect_evidence_r02c2.sh   (under gpurun, one GPU; outputs in gpurun_out/r02c/)
O=gpurun_out/r02c; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -k "iteration_count or config3" 2>&1 | tail -15 > $O/test_gpu_new.log
timeout 400 bash scripts/collect_counters_r02c.sh > $O/counters.log 2>&1
mkdir -p profiles/r02c_counters && cp gpurun_out/counters/robots_256_2000x56.csv gpurun_out/counters/obstacles_dense_16384x56.csv profiles/r02c_counters/ 2>/dev/null
python scripts/counters_to_json.py r02c_counters > $O/counters_json.log 2>&1
cp profiles/ncu_kernel_counters.json $O/ncu_kernel_counters.json
timeout 600 python bench.py > $O/bench_default_1gpu.json 2> $O/bench_default_1gpu.err
cat $O/test_gpu_new.log; tail -n 3 $O/counters.log $O/bench_default_1gpu.err; ls -la gpurun_out/counters

set -u
OUT=gpurun_out
timeout 60 ./tools/cu/umma_rate_probe.bin > $OUT/r2e_umma_rate.txt 2>&1; echo "rate probe: $?"
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2e_chain_time.txt 2>&1; echo "chain time: $?"
timeout 120 python tools/chain_time.py > $OUT/r2e_chain_time_clean.txt 2>&1; echo "chain clean: $?"
timeout 600 python -m pytest tests/test_gpu_entropy.py tests/test_gpu_contra.py tests/test_gpu_step.py -q > $OUT/r2e_pytest.log 2>&1; echo "pytest: $?"
timeout 400 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2e_bench_n1.json 2>$OUT/r2e_bench_n1.err; echo "bench: $?"
cat $OUT/r2e_umma_rate.txt; tail -4 $OUT/r2e_chain_time.txt; tail -3 $OUT/r2e_pytest.log

set -u
# State-of-the-world call (round 2, second session): everything re-measured on one box, outputs kept under gpurun_out/r2g_*
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2g_smi.txt 2>&1
timeout 120 ./tools/cu/tc_selftest.bin all  > $OUT/r2g_selftest_all.txt 2>&1; echo "selftest all: $?"
timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2g_selftest_perf_v2.txt 2>&1; echo "perf v2: $?"
U2PL_CONV_V=1 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2g_selftest_perf_v1.txt 2>&1; echo "perf v1: $?"
timeout 60 ./tools/cu/umma_rate_probe.bin > $OUT/r2g_umma_rate.txt 2>&1; echo "rate probe: $?"
timeout 900 python -m pytest tests -m gpu -q > $OUT/r2g_pytest_gpu.log 2>&1; echo "pytest all: $?"
timeout 120 python tools/chain_time.py > $OUT/r2g_chain_time.txt 2>&1; echo "chain: $?"
C=19 timeout 120 python tools/chain_time.py > $OUT/r2g_chain_time_c19.txt 2>&1; echo "chain c19: $?"
timeout 120 python tools/contra_bench.py > $OUT/r2g_contra_bench.txt 2>&1; echo "contra_bench: $?"
timeout 200 python tools/conv_bench.py > $OUT/r2g_conv_bench.jsonl 2>$OUT/r2g_conv_bench.err; echo "conv_bench: $?"
timeout 600 python bench.py --steps 10 --warmup 3 --phases > $OUT/r2g_bench_n1.json 2>$OUT/r2g_bench_n1.err; echo "bench: $?"
timeout 300 python tools/step_profile.py > $OUT/r2g_step_profile.txt 2>$OUT/r2g_step_profile.err; echo "profile: $?"
tail -n 15 $OUT/r2g_pytest_gpu.log; cat $OUT/r2g_chain_time.txt | tail -2; tail -3 $OUT/r2g_contra_bench.txt

set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 120 ./tools/cu/tc_selftest.bin all > $OUT/r2h_selftest_all.txt 2>&1; echo "selftest all (flat, 4 stages): $?"
U2PL_CONV_STAGES=3 timeout 120 ./tools/cu/tc_selftest.bin conv > $OUT/r2h_selftest_conv_s3.txt 2>&1; echo "selftest conv (3 stages/128B rows): $?"
U2PL_CONV_STAGES=3 timeout 120 ./tools/cu/tc_selftest.bin stats > $OUT/r2h_selftest_stats_s3.txt 2>&1; echo "selftest stats (3 stages): $?"
timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2h_perf_flat_s4.txt 2>&1; echo "perf flat s4: $?"
U2PL_CONV_STAGES=3 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2h_perf_flat_s3.txt 2>&1; echo "perf flat s3: $?"
U2PL_CONV_V=1 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2h_perf_v1.txt 2>&1; echo "perf v1: $?"
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2h_chain_time.txt 2>&1; echo "chain timing: $?"
timeout 900 python -m pytest tests -m gpu -q > $OUT/r2h_pytest_gpu.log 2>&1; echo "pytest all: $?"
U2PL_ENTROPY_CHAIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2h_bench_n1_nochain.json 2>$OUT/r2h_bench_n1_nochain.err; echo "bench: $?"
U2PL_ENTROPY_CHAIN=0 timeout 300 python tools/step_profile.py > $OUT/r2h_step_profile.txt 2>$OUT/r2h_step_profile.err; echo "profile: $?"
cat $OUT/r2h_selftest_all.txt | grep -v OK; grep -v OK $OUT/r2h_selftest_conv_s3.txt $OUT/r2h_selftest_stats_s3.txt; cat $OUT/r2h_perf_flat_s4.txt $OUT/r2h_perf_flat_s3.txt; tail -n 8 $OUT/r2h_pytest_gpu.log; tail -3 $OUT/r2h_chain_time.txt

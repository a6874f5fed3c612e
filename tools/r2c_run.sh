set -u
OUT=gpurun_out
timeout 120 ./tools/cu/tc_selftest.bin conv  > $OUT/r2c_selftest_conv.txt 2>&1; echo "selftest conv: $?"
timeout 120 ./tools/cu/tc_selftest.bin stats > $OUT/r2c_selftest_stats.txt 2>&1; echo "selftest stats: $?"
timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2c_selftest_perf_v2.txt 2>&1; echo "perf v2: $?"
U2PL_CONV_V=1 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2c_selftest_perf_v1.txt 2>&1; echo "perf v1: $?"
timeout 300 python -m pytest tests/test_gpu_entropy.py -x -q > $OUT/r2c_pytest_entropy.log 2>&1; echo "pytest entropy: $?"
timeout 600 python -m pytest tests -m gpu -q > $OUT/r2c_pytest_gpu.log 2>&1; echo "pytest all: $?"
timeout 200 python tools/conv_bench.py > $OUT/r2c_conv_bench.jsonl 2>$OUT/r2c_conv_bench.err; echo "conv_bench: $?"
timeout 400 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2c_bench_n1.json 2>$OUT/r2c_bench_n1.err; echo "bench: $?"
U2PL_TC_CONV=1 timeout 400 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2c_bench_n1_tcconv1.json 2>$OUT/r2c_bench_n1_tcconv1.err; echo "bench tcconv=1: $?"
U2PL_TC_CONV=1 U2PL_TC_TRAIN=1 timeout 400 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2c_bench_n1_tctrain.json 2>$OUT/r2c_bench_n1_tctrain.err; echo "bench tctrain: $?"
timeout 300 python tools/step_profile.py > $OUT/r2c_step_profile.txt 2>$OUT/r2c_step_profile.err; echo "profile: $?"
tail -n 4 $OUT/r2c_pytest_gpu.log $OUT/r2c_pytest_entropy.log; cat $OUT/r2c_selftest_conv.txt $OUT/r2c_selftest_stats.txt | tail -30

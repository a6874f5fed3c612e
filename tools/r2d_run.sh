set -u
OUT=gpurun_out
U2PL_CONV_DEBUG=1 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2d_selftest_perf_v2.txt 2>&1; echo "perf v2: $?"
U2PL_CONV_DEBUG=1 U2PL_CONV2_CLUSTERS=74 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2d_selftest_perf_v2_c74.txt 2>&1; echo "perf v2 c74: $?"
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2d_chain_time.txt 2>&1; echo "chain time: $?"
timeout 120 python tools/chain_time.py > $OUT/r2d_chain_time_clean.txt 2>&1; echo "chain time clean: $?"
timeout 600 python -m pytest tests -m gpu -q > $OUT/r2d_pytest_gpu.log 2>&1; echo "pytest all: $?"
timeout 400 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2d_bench_n1.json 2>$OUT/r2d_bench_n1.err; echo "bench: $?"
U2PL_FUSED_UP=0 timeout 400 python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2d_bench_n1_nofusedup.json 2>$OUT/r2d_bench_n1_nofusedup.err; echo "bench no fused up: $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 1 -o $OUT/r02_conv_tc2_v0 ./tools/cu/tc_selftest.bin perf > $OUT/r2d_ncu_conv.log 2>&1; echo "ncu conv: $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:entropy_chain -s 2 -c 1 -o $OUT/r02_entropy_chain_v0 python tools/chain_time.py > $OUT/r2d_ncu_chain.log 2>&1; echo "ncu chain: $?"
tail -n 6 $OUT/r2d_pytest_gpu.log; cat $OUT/r2d_chain_time.txt | tail -8; grep -h "conv_tc2\|layer3.conv2\|8192" $OUT/r2d_selftest_perf_v2.txt $OUT/r2d_selftest_perf_v2_c74.txt

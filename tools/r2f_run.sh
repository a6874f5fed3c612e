set -u
OUT=gpurun_out
U2PL_CONV2_TRACE=1 timeout 120 ./tools/cu/tc_selftest.bin perf > $OUT/r2f_conv_trace.txt 2>&1; echo "conv trace: $?"
timeout 300 python -m pytest tests/test_gpu_entropy.py -q -x > $OUT/r2f_pytest_entropy.log 2>&1; echo "pytest entropy: $?"
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2f_chain_time.txt 2>&1; echo "chain time: $?"
timeout 120 python tools/chain_time.py > $OUT/r2f_chain_time_clean.txt 2>&1; echo "chain clean: $?"
C=19 timeout 120 python tools/chain_time.py > $OUT/r2f_chain_time_c19.txt 2>&1; echo "chain c19: $?"
tail -3 $OUT/r2f_pytest_entropy.log; tail -4 $OUT/r2f_chain_time.txt; cat $OUT/r2f_chain_time_clean.txt $OUT/r2f_chain_time_c19.txt | tail -2; head -60 $OUT/r2f_conv_trace.txt

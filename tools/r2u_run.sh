set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 run 60 "chain dbg1" python tools/chain_time.py > $OUT/r2u_chain_time_dbg1.txt 2>&1
grep "entropy_chain. " $OUT/r2u_chain_time_dbg1.txt | tail -3
run 60 "chain clean" python tools/chain_time.py > $OUT/r2u_chain_time_clean.txt 2>&1
grep fused $OUT/r2u_chain_time_clean.txt
run 200 "pytest entropy contra" python -m pytest tests/test_gpu_entropy.py tests/test_gpu_contra.py -q > $OUT/r2u_pytest.log 2>&1
tail -2 $OUT/r2u_pytest.log

set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 run 60 "chain dbg1" python tools/chain_time.py > $OUT/r2n_chain_time_dbg1.txt 2>&1
grep entropy_chain $OUT/r2n_chain_time_dbg1.txt | tail -2
run 60 "chain clean" python tools/chain_time.py > $OUT/r2n_chain_time_clean.txt 2>&1
C=19 run 60 "chain c19" python tools/chain_time.py > $OUT/r2n_chain_time_c19.txt 2>&1
cat $OUT/r2n_chain_time_clean.txt $OUT/r2n_chain_time_c19.txt | grep fused
run 120 "pytest entropy" python -m pytest tests/test_gpu_entropy.py tests/test_gpu_step.py tests/test_gpu_dropin_api.py -q -x > $OUT/r2n_pytest_entropy.log 2>&1
tail -2 $OUT/r2n_pytest_entropy.log
U2PL_TC_CONV=1 run 200 "phase profile tcconv=1" python tools/phase_profile.py > $OUT/r2n_phase_profile_tcconv1.txt 2>&1
run 200 "phase profile default" python tools/phase_profile.py > $OUT/r2n_phase_profile.txt 2>&1

set -u
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 timeout 120 python tools/chain_time.py > $OUT/r2k_chain_time_dbg1.txt 2>&1; echo "chain dbg1: $?"
U2PL_CHAIN_TIMING=3 timeout 120 python tools/chain_time.py > $OUT/r2k_chain_time_dbg3.txt 2>&1; echo "chain dbg3 (no hist): $?"
grep entropy_chain $OUT/r2k_chain_time_dbg1.txt | tail -4; grep entropy_chain $OUT/r2k_chain_time_dbg3.txt | tail -4

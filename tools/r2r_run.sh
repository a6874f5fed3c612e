set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 run 60 "chain dbg1" python tools/chain_time.py > $OUT/r2r_chain_time_dbg1.txt 2>&1
grep entropy_chain $OUT/r2r_chain_time_dbg1.txt | tail -2
run 60 "chain clean" python tools/chain_time.py > $OUT/r2r_chain_time_clean.txt 2>&1
C=19 run 60 "chain c19" python tools/chain_time.py > $OUT/r2r_chain_time_c19.txt 2>&1
cat $OUT/r2r_chain_time_clean.txt $OUT/r2r_chain_time_c19.txt | grep fused
run 300 "pytest entropy+step" python -m pytest tests/test_gpu_entropy.py tests/test_gpu_step.py tests/test_gpu_dropin_api.py tests/test_gpu_fused.py -q > $OUT/r2r_pytest.log 2>&1
tail -4 $OUT/r2r_pytest.log
run 300 "bench" python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2r_bench.json 2>$OUT/r2r_bench.err
python - <<'PY'
import json
for f in ['r2r_bench']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').readline()); print(f, d['ms_per_step'], d['value'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY

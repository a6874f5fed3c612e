set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
U2PL_CHAIN_TIMING=1 run 60 "chain dbg1" python tools/chain_time.py > $OUT/r2q_chain_time_dbg1.txt 2>&1
grep entropy_chain $OUT/r2q_chain_time_dbg1.txt | tail -2
run 60 "chain clean" python tools/chain_time.py > $OUT/r2q_chain_time_clean.txt 2>&1
C=19 run 60 "chain c19" python tools/chain_time.py > $OUT/r2q_chain_time_c19.txt 2>&1
cat $OUT/r2q_chain_time_clean.txt $OUT/r2q_chain_time_c19.txt | grep fused
run 600 "pytest gpu" python -m pytest tests -m gpu -q > $OUT/r2q_pytest_gpu.log 2>&1
tail -4 $OUT/r2q_pytest_gpu.log
run 400 "bench" python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2q_bench.json 2>$OUT/r2q_bench.err
run 600 "bench fp32" python bench.py --fp32 --steps 5 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2q_bench_fp32.json 2>$OUT/r2q_bench_fp32.err
python - <<'PY'
import json
for f in ['r2q_bench','r2q_bench_fp32']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').readline()); print(f, d['ms_per_step'], d['value'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY

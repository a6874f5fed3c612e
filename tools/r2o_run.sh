set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 600 "pytest gpu" python -m pytest tests -m gpu -q -x > $OUT/r2o_pytest_gpu.log 2>&1
tail -4 $OUT/r2o_pytest_gpu.log
run 400 "bench" python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2o_bench.json 2>$OUT/r2o_bench.err
U2PL_BIAS_FOLD=0 run 400 "bench nobiasfold" python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2o_bench_nobiasfold.json 2>$OUT/r2o_bench_nobiasfold.err
run 400 "bench c2" python bench.py --workload c2 --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2o_bench_c2.json 2>$OUT/r2o_bench_c2.err
run 200 "ncu chain" ncu --set full --clock-control none --import-source on -k regex:entropy_chain -s 2 -c 1 -o $OUT/r02_entropy_chain_v5 python tools/chain_time.py > $OUT/r2o_ncu_chain.log 2>&1
python - <<'PY'
import json
for f in ['r2o_bench','r2o_bench_nobiasfold','r2o_bench_c2']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').readline()); print(f, d['ms_per_step'], d['value'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
PY

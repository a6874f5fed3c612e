set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 300 "smoke" python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2x_smoke.log 2>&1
tail -3 $OUT/r2x_smoke.log
run 900 "pytest gpu" python -m pytest tests -m gpu -q > $OUT/r2x_pytest_gpu.log 2>&1
tail -3 $OUT/r2x_pytest_gpu.log
run 600 "bench" python bench.py --steps 10 --warmup 3 --phases --no-cpu-baseline --no-eager-baseline > $OUT/r2x_bench.json 2>$OUT/r2x_bench.err
python - <<'PY'
import json
for f in ['r2x_bench']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').readline()); print(f, d['ms_per_step'], d['value'], d['phases_ms'], d['losses'], d['roofline']['us_per_call'], d['roofline']['frac'], d['roofline']['traffic'])
    except Exception as e: print(f, 'ERR', e)
PY

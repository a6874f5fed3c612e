set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 300 "pytest 2gpu tests" python -m pytest tests/test_gpu_sharded_bank.py tests/test_gpu_syncbn.py -q > $OUT/r2w_pytest_2gpu.log 2>&1
tail -4 $OUT/r2w_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run 400 "bench n4" $TR --master-port 29521 bench.py --gpus 4 --steps 8 --warmup 3 --phases > $OUT/r2w_bench_n4.json 2>$OUT/r2w_bench_n4.err
python - <<'PY'
import json
for f in ['r2w_bench_n4']:
    try:
        for line in open(f'gpurun_out/{f}.json'):
            if line.startswith('{'):
                d=json.loads(line); print(f, d['ms_per_step'], d['value'], d['phases_ms'], d['losses'], d['config']['bank']); break
        else: print(f, 'no json'); print(open(f'gpurun_out/{f}.err').read()[-1500:])
    except Exception as e: print(f, 'ERR', e)
PY

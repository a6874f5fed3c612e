set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 300 "pytest fused+model+dropin" python -m pytest tests/test_gpu_fused.py tests/test_gpu_dropin_api.py tests/test_gpu_conv_tc.py -q > $OUT/r2zz_pytest.log 2>&1
tail -3 $OUT/r2zz_pytest.log

set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 200 "pytest c2 full size" python -m pytest tests/test_gpu_entropy.py::test_full_size_c2_properties -q > $OUT/r2y_pytest_c2.log 2>&1
tail -15 $OUT/r2y_pytest_c2.log | cut -c1-300

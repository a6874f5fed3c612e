set -u
. tools/gpu_safe.sh
OUT=gpurun_out
mkdir -p $OUT
run 100 "ncu conv3+res" ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 18 -c 1 -o $OUT/r02_conv_flat_final_l3conv3_res ./tools/cu/tc_selftest.bin perf 2 > $OUT/r2z_ncu1.log 2>&1
run 100 "ncu l4conv2" ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 5 -c 1 -o $OUT/r02_conv_flat_final_l4conv2 ./tools/cu/tc_selftest.bin perf 3 > $OUT/r2z_ncu2.log 2>&1

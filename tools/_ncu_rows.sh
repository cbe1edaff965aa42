export FOCAL_B200_LIB=$PWD/focal_b200/libfocal_b200_a_pf2.so
ncu --set full --import-source on --clock-control none -k regex:"_v3_kernel" -s 4 -c 2 -o gpurun_out/r2_rows_v3b -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustain-seconds 0 > gpurun_out/b_ncu2.log 2>&1
tail -2 gpurun_out/b_ncu2.log

set -x
HBN_FP_G=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_astar_g -s 3 -c 1 -o gpurun_out/r1f_astar_g8 -f python bench.py --steps 1 --warmup 3 --queries 100000 --no-cpu-baseline > gpurun_out/r1f_ncu_g8.log 2>&1
tail -2 gpurun_out/r1f_ncu_g8.log | cut -c1-300
HBN_FP_G=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_astar_g -s 3 -c 1 -o gpurun_out/r1f_astar_g32 -f python bench.py --steps 1 --warmup 3 --queries 100000 --no-cpu-baseline > gpurun_out/r1f_ncu_g32.log 2>&1
tail -2 gpurun_out/r1f_ncu_g32.log | cut -c1-300

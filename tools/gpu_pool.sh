timeout 300 python -m pytest tests/test_vae_gpu.py tests/test_collate.py -m gpu -x -q -k "csr or layer or golden or collate or prefetch" 2>&1 | tail -5
for v in "" 4,32,8 4,16,8 4,8,8 4,32,4 4,32,16 4,16,16 2,32 1,32; do
  SLN_POOL=$v timeout 120 python tools/bench_pool.py 512 8192 2>&1 | tail -2
done > gpurun_out/pool_sweep.txt 2>&1
for v in 4,1,8 4,1,16 4,1,4 4,2,16 2,1 1,1; do
  SLN_POOL=$v timeout 120 python tools/bench_pool.py 64 2>&1 | tail -1
done >> gpurun_out/pool_sweep.txt 2>&1
cat gpurun_out/pool_sweep.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pool_fwd -s 3 -c 1 -o gpurun_out/prof_pool -f python tools/bench_pool.py 8192 > gpurun_out/prof_pool.log 2>&1
tail -2 gpurun_out/prof_pool.log

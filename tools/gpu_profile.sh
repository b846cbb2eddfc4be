# Regenerates every capture that profiles/ summarises (run through gpurun; then `python tools/summarize_profiles.py <tag>` here).
# gpurun copies back at most 64 MiB: every .ncu-rep is converted to its raw-page CSV on the box and big reports are dropped.
NCU_L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
NCU_F="ncu --set full --clock-control none -f"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -8 gpurun_out/smoke.txt
if [ -z "$SKIP_BENCH" ]; then
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_vae.json 2> gpurun_out/bench_vae.err
timeout 600 python bench.py --workload render --steps 50 --warmup 5 > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err
timeout 600 python bench.py --workload spade --steps 5 --warmup 3 > gpurun_out/bench_spade.json 2> gpurun_out/bench_spade.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_vae_reference.json 2> gpurun_out/bench_vae_reference.err
fi
timeout 300 $NCU_L --log-file gpurun_out/launches_vae.csv python tools/prof_step.py 3 > /dev/null 2>&1
timeout 300 $NCU_L --cache-control none --log-file gpurun_out/launches_vae_warm.csv python tools/prof_step.py 3 > /dev/null 2>&1
timeout 400 $NCU_L --cache-control none -c 1400 --log-file gpurun_out/launches_render_warm.csv python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 $NCU_L -c 1400 --log-file gpurun_out/launches_render.csv python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 $NCU_L --log-file gpurun_out/launches_spade.csv python tools/prof_spade.py 16 > /dev/null 2>&1
timeout 300 $NCU_F -k regex:tc_gemm -s 171 -c 4 -o gpurun_out/prof_tc_vae python tools/prof_step.py 2 > /dev/null 2>&1
timeout 300 $NCU_F --import-source on -k regex:tc_gemm -s 47 -c 2 -o gpurun_out/prof_tc_spade python tools/prof_spade.py 16 > /dev/null 2>&1
timeout 300 $NCU_F --import-source on -k regex:k_raster_tiles -s 3 -c 1 -o gpurun_out/prof_raster python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 $NCU_F -k regex:k_pool_fwd -s 3 -c 1 -o gpurun_out/prof_pool python tools/bench_pool.py 8192 > /dev/null 2>&1
timeout 300 $NCU_F -k regex:k_pool_fwd_node -s 12 -c 1 -o gpurun_out/prof_pool64 python tools/prof_step.py 2 > /dev/null 2>&1
for r in gpurun_out/*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null
  if [ $(stat -c %s $r) -gt 16000000 ]; then rm -f $r; fi
done
ls -la gpurun_out; du -sm gpurun_out
for f in vae render spade vae_reference; do tail -c 400 gpurun_out/bench_$f.json; echo; tail -c 300 gpurun_out/bench_$f.err; done

# end-of-round check: full GPU suite + smoke, and the render launch lists of the final code
NCU_L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -6 gpurun_out/smoke.txt
timeout 400 $NCU_L --cache-control none -c 900 --log-file gpurun_out/launches_render_warm.csv python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 $NCU_L -c 900 --log-file gpurun_out/launches_render.csv python bench.py --workload render --no-graph --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python bench.py --workload render --steps 50 --warmup 5 > gpurun_out/bench_render.json 2> gpurun_out/bench_render.err
tail -c 300 gpurun_out/bench_render.json

set -x
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_fullsize.log 2>&1; echo "fullsize rc=$?"
tail -15 gpurun_out/pytest_fullsize.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_g2p|k_p2g|k_sort|k_scan|k_mark|k_fill|k_extend|k_cg_build|k_pressure_patch|k_prev" -s 16 -c 16 -o gpurun_out/prof_stages_4096 python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 32 > gpurun_out/ncu_stages.log 2>&1; echo "ncu stages rc=$?"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc=$?"
cat gpurun_out/bench_default.json
timeout 600 python bench.py --workload sl1024 > gpurun_out/bench_sl1024.json 2> gpurun_out/bench_sl1024.err; echo "bench sl1024 rc=$?"
cat gpurun_out/bench_sl1024.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc=$?"
cat gpurun_out/bench_reference.json

# round 1, call v (1 GPU): tests + default bench after the division fast path in the set-up / patch kernels; ncu of the final stage kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_v.json'));print(d['ms_per_step'], d['stage_ms_per_step']); print(d['stage_roofline']); print(d['roofline']['avg_iteration_us'], d['e2e'])"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_g2p|k_p2g|k_sort|k_scan|k_mark|k_fill|k_extend|k_cg_build|k_pressure_patch|k_prev|k_enforce" -s 18 -c 22 -o gpurun_out/prof_stages_4096_v python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 32 > gpurun_out/ncu_stages_v.log 2>&1; echo "ncu stages rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_picflip4096_v.csv python bench.py --workload picflip4096 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --cg-cap 64 > gpurun_out/ncu_launches_v.log 2>&1; echo "ncu launches rc=$?"

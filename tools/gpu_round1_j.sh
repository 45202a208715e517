set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pressure or config0" > gpurun_out/pytest_pressure.log 2>&1; rc=$?; echo "pressure rc=$rc"
tail -5 gpurun_out/pytest_pressure.log
if [ $rc -ne 0 ]; then exit 1; fi
for wl in cg1024 cg4096; do
for pdl in 1 0; do
FSB_CG_PDL=$pdl timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_pdl$pdl.json 2> gpurun_out/bench_${wl}_pdl$pdl.err; echo "rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_${wl}_pdl$pdl.json'));print('$wl pdl$pdl', d['roofline']['avg_iteration_us'], d['roofline']['frac'], d['cg_iters_per_step'], d['ms_per_step'])"
done
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log

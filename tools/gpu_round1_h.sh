set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "pressure or config0" > gpurun_out/pytest_pressure.log 2>&1; rc=$?; echo "pressure rc=$rc"
tail -5 gpurun_out/pytest_pressure.log
if [ $rc -ne 0 ]; then exit 1; fi
for wl in cg1024 cg4096; do
for mode in fused 2k; do
FSB_CG_MODE=$mode timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_n1_$mode.json 2> gpurun_out/bench_${wl}_n1_$mode.err; echo "rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_${wl}_n1_$mode.json'));print('$wl n1 $mode', d['roofline']['avg_iteration_us'], d['roofline']['frac'], d['cg_iters_per_step'], d['ms_per_step'])"
done
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multi_gpu_cg_check.py --grid 512 > gpurun_out/mgpu_512.log 2>&1; echo "mgpu512 rc=$?"
grep "^{" gpurun_out/mgpu_512.log
for wl in cg1024 cg4096; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_${wl}_n2_fused.json 2> gpurun_out/bench_${wl}_n2_fused.err; echo "rc=$?"
grep "^{" gpurun_out/bench_${wl}_n2_fused.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$wl n2 fused', d['roofline']['avg_iteration_us'], d['roofline']['frac'], d['cg_iters_per_step'], d['ms_per_step'])"
done

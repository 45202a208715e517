# strong scaling of the CG-only workload (BASELINE configs[3]) on one 8-GPU box: N = 1, 2, 4, 8
set -x
WL=${1:-cg8192}
for n in ${SCALE_NS:-1 2 4 8}; do
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/scale_${WL}_n$n.json 2> gpurun_out/scale_${WL}_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/scale_${WL}_n$n.json 2> gpurun_out/scale_${WL}_n$n.err
  fi
  echo "rc=$?"
  grep "^{" gpurun_out/scale_${WL}_n$n.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$WL n$n', 'us/iter', d['roofline']['avg_iteration_us'], 'frac', d['roofline']['frac'], 'iters', d['cg_iters_per_step'], 'ms', d['ms_per_step'], 'iters/s', d['cg_iters_per_s'])"
done

# round 1, call w (4 GPUs): strong scaling of the 8192^2 CG at N = 1, 2, 4 and the default workload at N = 4 (what the driver's scaling run does)
set -x
mkdir -p gpurun_out
SCALE_NS="1 2 4" bash tools/gpu_scale.sh cg8192 2>&1 | grep -v "^+" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_4gpu_picflip4096.json 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"
head -c 300 gpurun_out/bench_4gpu_picflip4096.json; echo
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_4gpu_picflip4096.json') if l.startswith('{')][-1]);print(d['ms_per_step'], d['cg_iters_per_step'], d['roofline']['avg_iteration_us'], d['roofline']['kernel'][:40], d['e2e'])"

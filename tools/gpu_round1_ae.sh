# round 1, call ae (1 GPU): in-cell order by global id, particle slabs (in-process and over gloo on one GPU), full suite
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log

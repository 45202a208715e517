# round 1, call u (1 GPU): parity tests after the state-file fix
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"

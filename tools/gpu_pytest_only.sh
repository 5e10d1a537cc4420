set -u
O=gpurun_out/r3d; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -12 $O/pytest_gpu.log; grep -n "identical to the oracle" $O/pytest_gpu.log | head

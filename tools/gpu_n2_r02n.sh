timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "routed or group or sliced" > gpurun_out/pytest_routed_r02n.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_routed_r02n.log
QUICK=1 bash tools/gpu_multi.sh r02n 2 "c5:routed c5:routed:--exchange,nccl::nccl"
bash tools/gpu_multi.sh r02nv 2 "c2:routed"

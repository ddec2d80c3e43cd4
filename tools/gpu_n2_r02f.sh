timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "routed or group or sliced" > gpurun_out/pytest_routed_r02f.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_routed_r02f.log
bash tools/gpu_multi.sh r02f 2 "c5:routed c5:routed:--no-pipeline c5:reads"

set -x
K=10 timeout 90 python tools/_gpu_dbg.py > gpurun_out/smoke9.log 2>&1 || { tail -5 gpurun_out/smoke9.log; echo SMOKE_FAILED; exit 1; }
tail -2 gpurun_out/smoke9.log
export B200DOCK_TEST_KERNELS=10
timeout 300 python -m pytest tests/test_pose_init.py -m gpu -q 2>&1 | grep -v Warning | tail -30 > gpurun_out/t9a.log; tail -3 gpurun_out/t9a.log
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/t9.log; tail -4 gpurun_out/t9.log

set -x
timeout 300 python -m pytest tests/test_imgproc_gpu.py -x -q -m gpu > gpurun_out/s10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s10_pytest.log
python tools/gpu_probe_hbm_stages.py > gpurun_out/s10_hbm_v2.log 2>&1
tail -3 gpurun_out/s10_pytest.log; cat gpurun_out/s10_hbm_v2.log

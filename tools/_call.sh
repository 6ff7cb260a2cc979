set -x
python -m pytest tests/test_gemm_gpu.py tests/test_trocr_gpu.py tests/test_craft_gpu.py -x -q -m gpu > gpurun_out/s3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s3_pytest.log
python tools/gpu_probe_gemm_sustained.py > gpurun_out/s3_gemm_sust.log 2>&1
python tools/gpu_probe_gemm_burst.py > gpurun_out/s3_gemm_burst.log 2>&1
python tools/gpu_probe_encoder.py > gpurun_out/s3_enc.log 2>&1
tail -3 gpurun_out/s3_pytest.log; cat gpurun_out/s3_gemm_sust.log gpurun_out/s3_enc.log

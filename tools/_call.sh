set -x
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_trocr_gpu.py -x -q -m gpu > gpurun_out/s6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s6_pytest.log
python tools/gpu_probe_gemm_sustained.py > gpurun_out/s6_gemm_sust.log 2>&1
python tools/gpu_probe_encoder.py > gpurun_out/s6_enc.log 2>&1
python bench.py > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
tail -3 gpurun_out/s6_pytest.log; cat gpurun_out/s6_gemm_sust.log gpurun_out/s6_enc.log; head -c 600 gpurun_out/s6_bench.json

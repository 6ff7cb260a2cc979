set -x
NCROPS=512 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_tc_persist' --launch-skip 2 -c 1 -f -o gpurun_out/s14_attn_pair python tools/gpu_probe_attn.py > gpurun_out/s14_ncu.log 2>&1
ls -la gpurun_out/s14*

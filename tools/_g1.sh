mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -2
{
SWEEP_CFGS="[dict()]" timeout 300 python tools/sweep.py 48 540 960 4 2>&1
SWEEP_CFGS="[dict()]" timeout 300 python tools/sweep.py 96 540 960 1 2>&1
SWEEP_CFGS="[dict()]" timeout 300 python tools/sweep.py 54 720 1280 1 2>&1
} > gpurun_out/sweep34.log 2>&1
cat gpurun_out/sweep34.log

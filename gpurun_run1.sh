mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py -q --tb=short 2>&1) > gpurun_out/t_kernels.log
(R4R_CONV_MODE=exact timeout 900 python -m pytest tests/test_gpu_models.py -q --tb=short 2>&1) > gpurun_out/t_models.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -20) > gpurun_out/t_smoke.log
tail -5 gpurun_out/t_kernels.log; tail -30 gpurun_out/t_models.log; tail gpurun_out/t_smoke.log

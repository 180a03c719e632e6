mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r2e_pytest.log
tail -12 gpurun_out/r2e_pytest.log

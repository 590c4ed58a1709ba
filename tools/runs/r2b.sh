mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 -k "full_cov or fullcov or full_covariance or moment_matching" > gpurun_out/r2b_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.txt
tail -40 gpurun_out/r2b_pytest.txt

# subset of the GPU tests (run under gpurun)
set -x
O=gpurun_out
python -m pytest tests -m gpu -q -x -k "coverage" > $O/pytest_cov.log 2>&1; tail -30 $O/pytest_cov.log

# K1 timing of the in-tree library + the GPU test suite (run under gpurun)
set -x
O=gpurun_out
rm -f $O/k1_ab.log
python scratch/k1_bench.py 2>&1 | grep -v Warn >> $O/k1_ab.log
python -m pytest tests -m gpu -q -x > $O/pytest_r2k.log 2>&1; tail -3 $O/pytest_r2k.log >> $O/k1_ab.log
cat $O/k1_ab.log

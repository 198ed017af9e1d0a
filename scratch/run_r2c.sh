set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_r2c.log 2>&1; tail -4 $O/pytest_r2c.log
for lib in scratch/variants/lib_pts1.so golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_pts4.so scratch/variants/lib_pts2mb9.so scratch/variants/lib_pts1mb9.so; do
  GOLEMFLAVOR_B200_LIB=$lib python scratch/k2_bench.py $O/k2_ref_r2.npy
done > $O/ab_r2c.log 2>&1
cat $O/ab_r2c.log
python scratch/ens_bench.py > $O/ens_r2c.log 2>&1; cat $O/ens_r2c.log
python scratch/sens_bench.py > $O/sens_r2c.log 2>&1; cat $O/sens_r2c.log
python scratch/call_latency.py > $O/lat_r2c.log 2>&1; cat $O/lat_r2c.log

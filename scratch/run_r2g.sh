set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_r2g.log 2>&1; tail -6 $O/pytest_r2g.log
for lib in scratch/variants/lib_lanes1.so golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_lanes2ilp5.so; do
  echo "== $lib"
  GOLEMFLAVOR_B200_LIB=$lib python scratch/ens_c3.py 2000 4096 20
  GOLEMFLAVOR_B200_LIB=$lib python scratch/ens_c3.py 2000 1024 20
  GOLEMFLAVOR_B200_LIB=$lib python scratch/ens_c3.py 2000 60 20
  GOLEMFLAVOR_B200_LIB=$lib python scratch/sens_bench.py
done > $O/lanes_r2g.log 2>&1
cat $O/lanes_r2g.log
python scratch/ens_bench.py 2>&1 | grep "C2" > $O/ens_r2g.log; cat $O/ens_r2g.log
python scratch/k2_bench.py $O/k2_ref_r2.npy

set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_r2h.log 2>&1; tail -6 $O/pytest_r2h.log
for lib in scratch/variants/lib_libndtri.so golemflavor_b200/lib/libgolemflavor_b200.so; do
  echo "== $lib"
  GOLEMFLAVOR_B200_LIB=$lib python scratch/scan_bench.py 1e9 x,texture,anarchic
  GOLEMFLAVOR_B200_LIB=$lib python scratch/evid_bench.py
done > $O/scan_r2h.log 2>&1
cat $O/scan_r2h.log
python scratch/ens_c3.py 2000 4096 20; python scratch/sens_bench.py

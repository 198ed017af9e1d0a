set -x
O=gpurun_out
for lib in scratch/variants/lib_head.so golemflavor_b200/lib/libgolemflavor_b200.so; do
  echo "== $lib"
  GOLEMFLAVOR_B200_LIB=$lib python scratch/k2_bench.py $O/k2_ref_r2.npy
  GOLEMFLAVOR_B200_LIB=$lib python scratch/scan_bench.py 1e9 texture,anarchic
  GOLEMFLAVOR_B200_LIB=$lib python scratch/evid_bench.py
done > $O/ab_rebuild.log 2>&1
cat $O/ab_rebuild.log

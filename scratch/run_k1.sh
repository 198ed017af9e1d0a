# A/B of the deferred-refinement bin loop (run under gpurun)
set -x
O=gpurun_out
rm -f $O/defer_ab.log
for v in golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_defer.so golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_defer.so; do
echo "== $v" >> $O/defer_ab.log
GOLEMFLAVOR_B200_LIB=$v python scratch/k2_bench.py $O/k2_ref_r2.npy 2>&1 | grep -v Warn >> $O/defer_ab.log
GOLEMFLAVOR_B200_LIB=$v python scratch/scan_bench.py 1e9 anarchic 2>&1 | grep "nb= 25" >> $O/defer_ab.log
GOLEMFLAVOR_B200_LIB=$v python scratch/evid_bench.py 2>&1 | grep evidence >> $O/defer_ab.log
done
cat $O/defer_ab.log

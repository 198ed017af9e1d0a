# K1 A/B of library variants (run under gpurun): timings + bit-compare into gpurun_out/k1_ab.log
set -x
O=gpurun_out
rm -f $O/k1_ref.npy $O/k1_ab.log
for v in golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_k1b18.so scratch/variants/lib_k1b20.so scratch/variants/lib_k1b24.so golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_k1b18.so; do
echo "== $v" >> $O/k1_ab.log
GOLEMFLAVOR_B200_LIB=$v python scratch/k1_bench.py $O/k1_ref.npy 2>&1 | grep -v Warn >> $O/k1_ab.log
done
cat $O/k1_ab.log

# K1 / K2 A/B of the in-tree library against saved variants (run under gpurun): timings + bit-compare into gpurun_out/k1_ab.log
set -x
O=gpurun_out
rm -f $O/k1_ref.npy $O/k1_ab.log
for v in scratch/variants/lib_k1p1.so golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_k1p1.so golemflavor_b200/lib/libgolemflavor_b200.so; do
echo "== $v" >> $O/k1_ab.log
GOLEMFLAVOR_B200_LIB=$v python scratch/k1_bench.py $O/k1_ref.npy 2>&1 | grep -v Warn >> $O/k1_ab.log
GOLEMFLAVOR_B200_LIB=$v python scratch/k2_bench.py $O/k2_ref_r2.npy 2>&1 | grep -v Warn >> $O/k1_ab.log
done
python scratch/ens_bench.py >> $O/k1_ab.log 2>&1
for m in texture anarchic; do python scratch/scan_bench.py 1e9 $m >> $O/k1_ab.log 2>&1; done
python -m pytest tests -m gpu -q -x > $O/pytest_r2i.log 2>&1; tail -3 $O/pytest_r2i.log >> $O/k1_ab.log
cat $O/k1_ab.log

# sampler A/B of library variants (run under gpurun)
set -x
O=gpurun_out
rm -f $O/ens_ab.log
for v in scratch/variants/lib_pch1.so golemflavor_b200/lib/libgolemflavor_b200.so scratch/variants/lib_pch1.so golemflavor_b200/lib/libgolemflavor_b200.so; do
echo "== $v" >> $O/ens_ab.log
GOLEMFLAVOR_B200_LIB=$v timeout 600 python scratch/ens_bench.py 2>&1 | grep "mode 3 nc  0" >> $O/ens_ab.log
GOLEMFLAVOR_B200_LIB=$v timeout 600 python scratch/ens_repeat.py 3 60 2000 4 >> $O/ens_ab.log 2>&1
done
timeout 600 python -m pytest tests -m gpu -q -x -k "sampler or mcmc or emcee or sens" > $O/pytest_ens.log 2>&1; tail -3 $O/pytest_ens.log >> $O/ens_ab.log
cat $O/ens_ab.log

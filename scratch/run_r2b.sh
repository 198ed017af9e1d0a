set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_r2b.log 2>&1; tail -5 $O/pytest_r2b.log
for n in 4096 60; do GOLEMFLAVOR_B200_LIB=scratch/variants/lib_ensprof.so python scratch/ens_c3.py 500 $n 20 ; done > $O/ens_stage.log 2>&1
cat $O/ens_stage.log
for lib in scratch/variants/lib_base.so golemflavor_b200/lib/libgolemflavor_b200.so; do
  GOLEMFLAVOR_B200_LIB=$lib python scratch/k2_bench.py $O/k2_ref_r2.npy
  GOLEMFLAVOR_B200_LIB=$lib python scratch/k1_bench.py
  GOLEMFLAVOR_B200_LIB=$lib python scratch/scan_bench.py 1e9 texture,anarchic
  GOLEMFLAVOR_B200_LIB=$lib python scratch/ens_c3.py 2000 4096 20
done > $O/ab_r2b.log 2>&1
cat $O/ab_r2b.log
python scratch/k2_bench.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_lnprob --launch-skip 4 -c 1 -f -o $O/prof_lnprob_r2b python scratch/k2_bench.py > $O/ncu_r2b.log 2>&1
ls -la $O/*.ncu-rep | tail -3

# Round-1 profile capture (run under gpurun): every ncu pass only after the same command exited 0 without ncu.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --scan-samples 100000000 --configs 0 --cpu-evals 0"
$B > $O/plain.log 2> $O/plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r01.csv $B > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lnprob --launch-skip 4 -c 1 -f -o $O/prof_lnprob_r01 $B > $O/ncu2.log 2>&1
python scratch/scan_bench.py 1e8 anarchic > $O/plain2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_hist --launch-skip 1 -c 1 -f -o $O/prof_hist_r01 python scratch/scan_bench.py 1e8 anarchic > $O/ncu3.log 2>&1
python scratch/k1_bench.py > $O/plain3.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_lnprob --launch-skip 4 -c 1 -f -o $O/prof_k1 python scratch/k1_bench.py > $O/ncu_k1.log 2>&1
python scratch/ens_one.py 1000 > $O/plain4.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_ensemble_cluster --launch-skip 1 -c 1 -f -o $O/prof_ens python scratch/ens_one.py 1000 > $O/ncu_ens.log 2>&1
ls -la $O

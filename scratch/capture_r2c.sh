# Final round-2 profile capture (run under gpurun): every ncu pass only after the same command exited 0 without ncu.
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --scan-samples 100000000 --configs 0 --cpu-evals 0 --sustain-s 0"
$B > $O/plain_r2c.log 2> $O/plain_r2c.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02c.csv $B > $O/ncu_l.log 2>&1
python scratch/k2_bench.py > $O/plain_k2.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_lnprob --launch-skip 4 -c 1 -f -o $O/prof_lnprob_r02c python scratch/k2_bench.py > $O/ncu_k2.log 2>&1
python scratch/scan_bench.py 1e8 anarchic > $O/plain_scan.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_hist --launch-skip 1 -c 1 -f -o $O/prof_hist_r02c python scratch/scan_bench.py 1e8 anarchic > $O/ncu_scan.log 2>&1
python scratch/k1_bench.py > $O/plain_k1.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_lnprob --launch-skip 4 -c 1 -f -o $O/prof_k1_r02c python scratch/k1_bench.py > $O/ncu_k1.log 2>&1
python scratch/ens_c3.py 300 4096 20 > $O/plain_ens.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_ensemble_cluster --launch-skip 1 -c 1 -f -o $O/prof_ens_r02c python scratch/ens_c3.py 300 4096 20 > $O/ncu_ens.log 2>&1
ls -la $O/*_r02c.ncu-rep; cat $O/plain_k2.log $O/plain_k1.log $O/plain_scan.log $O/plain_ens.log

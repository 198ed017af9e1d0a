set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_final.log 2>&1; tail -4 $O/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1; tail -2 $O/smoke_final.log
python scratch/k2_bench.py > $O/plain_k2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_lnprob --launch-skip 4 -c 1 -f -o $O/prof_lnprob_r02 python scratch/k2_bench.py > $O/ncu_k2.log 2>&1
python scratch/scan_bench.py 1e8 anarchic > $O/plain_scan.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_hist --launch-skip 1 -c 1 -f -o $O/prof_hist_r02 python scratch/scan_bench.py 1e8 anarchic > $O/ncu_scan.log 2>&1
B="python bench.py --steps 2 --warmup 3 --scan-samples 100000000 --configs 0 --cpu-evals 0 --sustain-s 0"
$B > $O/plain_r2.log 2> $O/plain_r2.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02.csv $B > $O/ncu_l.log 2>&1
python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err; tail -c 300 $O/bench_final.err

set -x
O=gpurun_out
python scratch/parity_scale.py $O/parity_scale_r02.json > $O/parity_scale_r02.log 2>&1; tail -5 $O/parity_scale_r02.log
python scratch/ens_bench.py 2>&1 | grep "C3" 

# run-to-run spread of the sampler shapes (run under gpurun)
set -x
O=gpurun_out
rm -f $O/ens_ab.log
for m in 1 3 1 3; do timeout 300 python scratch/ens_repeat.py $m 4096 1000 16 >> $O/ens_ab.log 2>&1; done
cat $O/ens_ab.log

# SM-only scans + scan tests + sharding invariance on the final build (run under gpurun)
set -x
O=gpurun_out
rm -f $O/scan_ab.log
python scratch/scan_bench.py 1e9 unitary,x >> $O/scan_ab.log 2>&1
python -m pytest tests -m gpu -q -x -k "scan or hist or cli or smoke" > $O/pytest_scan.log 2>&1; tail -3 $O/pytest_scan.log >> $O/scan_ab.log
cat $O/scan_ab.log

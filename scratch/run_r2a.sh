set -x
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_r2a.log 2>&1; tail -5 $O/pytest_r2a.log
python bench.py --steps 20 --warmup 5 > $O/bench_r2a.json 2> $O/bench_r2a.err; tail -c 600 $O/bench_r2a.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_r2a_ref.json 2> $O/bench_r2a_ref.err
GOLEMFLAVOR_B200_LIB=scratch/variants/lib_ensprof.so python scratch/ens_c3.py 500 4096 20 > $O/ens_c3_prof.log 2>&1
GOLEMFLAVOR_B200_LIB=scratch/variants/lib_ensprof.so python scratch/ens_c3.py 500 4096 1 >> $O/ens_c3_prof.log 2>&1
GOLEMFLAVOR_B200_LIB=scratch/variants/lib_ensprof.so python scratch/ens_c3.py 500 60 20 >> $O/ens_c3_prof.log 2>&1
cat $O/ens_c3_prof.log
python scratch/sens_bench.py > $O/sens_r2a.log 2>&1; cat $O/sens_r2a.log
python scratch/evid_bench.py > $O/evid_r2a.log 2>&1; cat $O/evid_r2a.log

# Round check (run under gpurun): GPU tests, smoke(), the driver's two bench commands.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_check.log 2>&1; tail -4 $O/pytest_check.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_check.log 2>&1; tail -2 $O/smoke_check.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_check_ref.json 2> $O/bench_check_ref.err
python bench.py --steps 20 --warmup 5 > $O/bench_check.json 2> $O/bench_check.err; tail -c 400 $O/bench_check.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err

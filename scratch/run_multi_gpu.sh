# Multi-GPU check (run under `gpurun --gpus N -- bash scratch/run_multi_gpu.sh N`): NCCL invariance check + the bench line at N ranks.
set -x
O=gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/dist_scan_check.py > $O/dist_scan_check_n$N.log 2>&1; echo rc=$? >> $O/dist_scan_check_n$N.log; grep -v "^W\|^\[W\|NCCL version" $O/dist_scan_check_n$N.log | tail -7
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_final_n$N.json 2> $O/bench_final_n$N.err; tail -c 300 $O/bench_final_n$N.err
python - <<PY
import json
d=json.load(open('$O/bench_final_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
print({k:v for k,v in d['config'].items() if k.startswith(('scan_','c1','c2','c3','c5','k1'))})
PY

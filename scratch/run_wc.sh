set -x
N=${1:-1}
O=gpurun_out
for wc in 0 1; do
  if [ "$N" = "1" ]; then E2E_WC=$wc python scratch/e2e_ranks.py 2>/dev/null | grep slots
  else E2E_WC=$wc python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 scratch/e2e_ranks.py 2>/dev/null | grep slots; fi
done > $O/e2e_wc_n$N.log; cat $O/e2e_wc_n$N.log

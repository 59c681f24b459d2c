set -u
O=gpurun_out/s27
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > $O/loc_n8.log 2>&1; echo "loc_n8 rc=$?" >> $O/status.txt )
cat $O/status.txt; grep '^{' $O/loc_n8.log | tail -1 | cut -c1-200

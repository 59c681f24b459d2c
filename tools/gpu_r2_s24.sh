set -u
O=gpurun_out/s24
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( GLOC_CSM_TIMING=1 timeout 600 $TR --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/loc_n8.log 2>&1; echo "loc_n8 rc=$?" >> $O/status.txt )
( timeout 900 $TR --master-port 29535 bench.py --gpus 8 --loc-rows 5000000 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg4_loc_5m_n8.log 2>&1; echo "cfg4_loc_5m_n8 rc=$?" >> $O/status.txt )
cat $O/status.txt; for f in loc_n8 cfg4_loc_5m_n8; do grep '^{' $O/$f.log | tail -1 | cut -c1-200; done

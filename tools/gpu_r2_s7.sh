set -u
O=gpurun_out/s7
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nvidia-smi -L > $O/gpus.txt 2>&1; nproc >> $O/gpus.txt; free -g >> $O/gpus.txt
( GLOC_CSM_TIMING=1 timeout 900 $TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/loc_n8.log 2>&1; echo "loc_n8 rc=$?" >> $O/status.txt )
( GLOC_SHARD_TIMING=1 timeout 900 $TR --master-port 29512 bench.py --gpus 8 --workload retrieval --rows 1000000 --queries 100000 --steps 5 --warmup 3 --no-cpu-baseline > $O/cfg3_n8.log 2>&1; echo "cfg3_n8 rc=$?" >> $O/status.txt )
( timeout 600 $TR --master-port 29513 bench.py --gpus 8 --workload stream --no-cpu-baseline > $O/stream_n8_1m.log 2>&1; echo "stream_n8_1m rc=$?" >> $O/status.txt )
( timeout 900 $TR --master-port 29514 bench.py --gpus 8 --workload stream --rows 5000000 --no-cpu-baseline > $O/stream_n8_5m.log 2>&1; echo "stream_n8_5m rc=$?" >> $O/status.txt )
( timeout 1200 $TR --master-port 29515 bench.py --gpus 8 --loc-rows 5000000 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg4_loc_5m_n8.log 2>&1; echo "cfg4_loc_5m_n8 rc=$?" >> $O/status.txt )
cat $O/status.txt; for f in loc_n8 cfg3_n8 stream_n8_1m stream_n8_5m cfg4_loc_5m_n8; do grep '^{' $O/$f.log | tail -1 | cut -c1-160; done

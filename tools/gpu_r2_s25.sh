set -u
O=gpurun_out/s25
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded_p2p.log 2>&1; echo "tests_sharded_p2p rc=$?" >> $O/status.txt )
( GLOC_SHARD_NO_P2P=1 GLOC_LOC_FULL_UPLOAD=1 timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded_nccl.log 2>&1; echo "tests_sharded_nccl rc=$?" >> $O/status.txt )
( timeout 900 $TR --master-port 29521 bench.py --gpus 2 --no-cpu-baseline > $O/loc_n2.log 2>&1; echo "loc_n2 rc=$?" >> $O/status.txt )
( GLOC_LOC_FULL_UPLOAD=1 timeout 900 $TR --master-port 29522 bench.py --gpus 2 --no-cpu-baseline > $O/loc_n2_fullupload.log 2>&1; echo "loc_n2_fullupload rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -15 $O/tests_sharded_p2p.log; for f in loc_n2 loc_n2_fullupload; do grep '^{' $O/$f.log | tail -1 | cut -c1-200; done

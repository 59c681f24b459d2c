set -u
O=gpurun_out/s23
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded_p2p.log 2>&1; echo "tests_sharded_p2p rc=$?" >> $O/status.txt )
( GLOC_SHARD_NO_P2P=1 timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded_nccl.log 2>&1; echo "tests_sharded_nccl rc=$?" >> $O/status.txt )
( timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py -x -q -m gpu > $O/tests_csm.log 2>&1; echo "tests_csm rc=$?" >> $O/status.txt )
( timeout 900 $TR --master-port 29521 bench.py --gpus 2 --no-cpu-baseline > $O/loc_n2.log 2>&1; echo "loc_n2 rc=$?" >> $O/status.txt )
( GLOC_BENCH_NO_SHARE=1 timeout 900 $TR --master-port 29522 bench.py --gpus 2 --no-cpu-baseline > $O/loc_n2_noshare.log 2>&1; echo "loc_n2_noshare rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -15 $O/tests_sharded_p2p.log; tail -3 $O/tests_csm.log; for f in loc_n2 loc_n2_noshare; do grep '^{' $O/$f.log | tail -1 | cut -c1-200; done

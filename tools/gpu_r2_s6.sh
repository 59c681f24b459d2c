set -u
O=gpurun_out/s6
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded_p2p.log 2>&1; echo "tests_sharded_p2p rc=$?" >> $O/status.txt )
( GLOC_SHARD_NO_P2P=1 timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded_nccl.log 2>&1; echo "tests_sharded_nccl rc=$?" >> $O/status.txt )
( timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py tests/test_host_cpp.py tests/test_driver.py -x -q -m gpu > $O/tests_csm.log 2>&1; echo "tests_csm rc=$?" >> $O/status.txt )
( GLOC_CSM_EXPAND=w timeout 900 python -m pytest tests/test_csm_gpu.py tests/test_localize_gpu.py -x -q -m gpu > $O/tests_csm_wide.log 2>&1; echo "tests_csm_wide rc=$?" >> $O/status.txt )
( GLOC_CSM_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/loc_timing.log 2>&1; echo "loc_timing rc=$?" >> $O/status.txt )
( GLOC_CSM_EXPAND=n GLOC_CSM_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/loc_timing_narrow.log 2>&1; echo "loc_timing_narrow rc=$?" >> $O/status.txt )
( timeout 600 $TR --master-port 29503 bench.py --gpus 2 --workload stream --no-cpu-baseline > $O/stream_n2_p2p.log 2>&1; echo "stream_n2_p2p rc=$?" >> $O/status.txt )
( GLOC_SHARD_NO_P2P=1 timeout 600 $TR --master-port 29504 bench.py --gpus 2 --workload stream --no-cpu-baseline > $O/stream_n2_nccl.log 2>&1; echo "stream_n2_nccl rc=$?" >> $O/status.txt )
( GLOC_SHARD_TIMING=1 timeout 900 $TR --master-port 29505 bench.py --gpus 2 --workload retrieval --rows 1000000 --queries 100000 --steps 3 --warmup 3 --no-cpu-baseline > $O/cfg3_n2_p2p.log 2>&1; echo "cfg3_n2_p2p rc=$?" >> $O/status.txt )
( timeout 900 $TR --master-port 29506 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/loc_n2.log 2>&1; echo "loc_n2 rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -3 $O/tests_sharded_p2p.log; grep "csm\]" $O/loc_timing.log | tail -2; grep "csm\]" $O/loc_timing_narrow.log | tail -2; grep "shard\]" $O/cfg3_n2_p2p.log | tail -3

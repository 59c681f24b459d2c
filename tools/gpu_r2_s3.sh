set -u
O=gpurun_out/s3
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi -L > $O/gpus.txt 2>&1
( timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded.log 2>&1; echo "tests_sharded rc=$?" >> $O/status.txt )
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/loc_n1.log 2>&1; echo "loc_n1 rc=$?" >> $O/status.txt )
( timeout 900 $TR --master-port 29501 bench.py --gpus 2 --steps 10 --warmup 3 > $O/loc_n2.log 2>&1; echo "loc_n2 rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --workload retrieval --rows 1000000 --queries 100000 --steps 5 --warmup 3 --no-cpu-baseline > $O/cfg3_n1.log 2>&1; echo "cfg3_n1 rc=$?" >> $O/status.txt )
( timeout 900 $TR --master-port 29502 bench.py --gpus 2 --workload retrieval --rows 1000000 --queries 100000 --steps 5 --warmup 3 --no-cpu-baseline > $O/cfg3_n2.log 2>&1; echo "cfg3_n2 rc=$?" >> $O/status.txt )
( timeout 600 $TR --master-port 29503 bench.py --gpus 2 --workload stream --no-cpu-baseline > $O/stream_n2.log 2>&1; echo "stream_n2 rc=$?" >> $O/status.txt )
( timeout 600 $TR --master-port 29504 bench.py --gpus 2 --workload retrieval --steps 10 --no-cpu-baseline > $O/cfg1_n2.log 2>&1; echo "cfg1_n2 rc=$?" >> $O/status.txt )
( timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_grid_store.py tests/test_csm_gpu.py tests/test_driver_network_gpu.py -x -q -m gpu > $O/tests_enc.log 2>&1; echo "tests_enc rc=$?" >> $O/status.txt )
cat $O/status.txt
tail -3 $O/tests_enc.log
tail -5 $O/tests_sharded.log

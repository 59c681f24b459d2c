set -u
O=gpurun_out/s34
mkdir -p $O
( timeout 300 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu > $O/tests_sharded_p2p.log 2>&1; echo "tests_sharded_p2p rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -15 $O/tests_sharded_p2p.log

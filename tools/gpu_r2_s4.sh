set -u
O=gpurun_out/s4
mkdir -p $O
( timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/status.txt )
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/loc_n1.log 2>&1; echo "loc_n1 rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --workload retrieval --rows 1000000 --queries 100000 --steps 5 --warmup 3 --no-cpu-baseline > $O/cfg3_n1.log 2>&1; echo "cfg3_n1 rc=$?" >> $O/status.txt )
( timeout 600 python bench.py --workload retrieval --steps 10 --no-cpu-baseline > $O/cfg1_n1.log 2>&1; echo "cfg1_n1 rc=$?" >> $O/status.txt )
( timeout 600 python bench.py --workload stream --no-cpu-baseline > $O/stream_n1.log 2>&1; echo "stream_n1 rc=$?" >> $O/status.txt )
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:enc_conv3x3_kernel -s 9 -c 3 -f -o $O/full_enc python bench.py --workload describe --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_enc.log 2>&1; echo "ncu_enc rc=$?" >> $O/status.txt )
cat $O/status.txt
tail -5 $O/tests_all.log

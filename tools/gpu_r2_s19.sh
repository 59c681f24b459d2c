set -u
O=gpurun_out/s19
mkdir -p $O
( timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests_all.log 2>&1; echo "tests_all rc=$?" >> $O/status.txt )
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/status.txt )
( timeout 900 python bench.py > $O/loc.log 2>&1; echo "loc rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/loc_ref.log 2>&1; echo "loc_ref rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --loc-policy first --no-cpu-baseline > $O/loc_first.log 2>&1; echo "loc_first rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --loc-queries 128 --no-cpu-baseline > $O/loc_q128.log 2>&1; echo "loc_q128 rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --workload retrieval --no-cpu-baseline > $O/retrieval.log 2>&1; echo "retrieval rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --workload stream --no-cpu-baseline > $O/stream.log 2>&1; echo "stream rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --workload verify --no-cpu-baseline > $O/verify.log 2>&1; echo "verify rc=$?" >> $O/status.txt )
( timeout 900 python bench.py --workload describe --no-cpu-baseline > $O/describe.log 2>&1; echo "describe rc=$?" >> $O/status.txt )
cat $O/status.txt; tail -3 $O/tests_all.log; for f in loc loc_ref loc_first loc_q128 retrieval stream verify describe; do grep '^{' $O/$f.log | tail -1 | cut -c1-200; done

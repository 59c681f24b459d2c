set -u
O=gpurun_out/s33
mkdir -p $O
( timeout 300 python bench.py --workload verify --no-cpu-baseline > $O/verify.log 2>&1; echo "verify rc=$?" >> $O/status.txt )
cat $O/status.txt; grep '^{' $O/verify.log | tail -1 | cut -c1-300; tail -3 $O/verify.log | cut -c1-300

#!/usr/bin/env bash
# Full validation on one B200: every GPU test, smoke, the default bench (both arms), streaming profiles.
set -u
OUT=gpurun_out
TAG=${1:-r01b}
mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/f_pytest.log 2>&1; echo "pytest rc=$?" > $OUT/f_status.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/f_smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/f_status.txt
timeout 900 python bench.py > $OUT/f_bench.json 2> $OUT/f_bench.err; echo "bench rc=$?" >> $OUT/f_status.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/f_bench_ref.json 2>> $OUT/f_bench.err; echo "bench-ref rc=$?" >> $OUT/f_status.txt
timeout 600 bash tools/profile_uvd.sh $TAG > $OUT/f_profile.log 2>&1; echo "profile rc=$?" >> $OUT/f_status.txt
cat $OUT/f_status.txt; tail -3 $OUT/f_pytest.log; tail -2 $OUT/f_smoke.log; head -c 400 $OUT/f_bench.json; echo; head -c 400 $OUT/f_bench_ref.json

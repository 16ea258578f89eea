#!/bin/bash
# Final GPU call of round 2: the whole -m gpu suite, smoke(), the default bench and the reference arm.
set -u
out=gpurun_out/r2final
mkdir -p $out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q -x > $out/tests.log 2>&1
echo "gpu suite exit $?" | tee -a $out/summary.txt
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1
echo "smoke exit $?" | tee -a $out/summary.txt
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench default exit $?" | tee -a $out/summary.txt
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
echo "bench reference exit $?" | tee -a $out/summary.txt
cp gpurun_out/parity_report.jsonl $out/ 2>/dev/null
tail -6 $out/tests.log | cut -c1-200
tail -2 $out/smoke.log
head -c 400 $out/bench_default.json; echo
cat $out/bench_reference.json | cut -c1-600

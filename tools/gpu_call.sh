#!/bin/bash
# Whole GPU suite on the final tree (per-test seeding, split-K heuristic), smoke().
set -u
out=gpurun_out/r2final3
mkdir -p $out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q > $out/tests.log 2>&1
echo "gpu suite exit $?" | tee -a $out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke exit $?" | tee -a $out/summary.txt
cp gpurun_out/parity_report.jsonl $out/ 2>/dev/null
grep -n "FAILED\|passed\|failed" $out/tests.log | tail -12 | cut -c1-250
tail -1 $out/smoke.log

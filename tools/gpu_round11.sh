#!/bin/bash
# correctness tooling: every kernel on small inputs under compute-sanitizer; the new cfg1 real-genome test
mkdir -p gpurun_out
timeout 300 python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_initcheck.log 2>&1; echo "initcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg1 or dropin" > gpurun_out/pytest_cfg1.log 2>&1; echo "cfg1 rc=$?"
tail -4 gpurun_out/sanitize_plain.log; for t in memcheck initcheck racecheck; do echo "== $t"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small" gpurun_out/sanitize_$t.log | tail -3; done; tail -3 gpurun_out/pytest_cfg1.log

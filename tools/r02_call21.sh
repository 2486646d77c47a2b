#!/bin/bash
# initcheck over smoke(): does any kernel of the level-6 / inflate path read device memory nobody wrote?
set -u
mkdir -p gpurun_out
timeout 75 compute-sanitizer --tool initcheck --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/initcheck_smoke.log 2>&1
echo "rc=$?"
grep -c "Uninitialized" gpurun_out/initcheck_smoke.log
grep -A6 "Uninitialized" gpurun_out/initcheck_smoke.log | grep -v "^--" | head -40
tail -4 gpurun_out/initcheck_smoke.log

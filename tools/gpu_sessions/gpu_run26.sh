#!/bin/bash
# last check of the committed tree: suite + smoke
set -x
mkdir -p gpurun_out
O=$PWD/gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/r02_pytest_gpu.log
tail -3 $O/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r02_smoke.log 2>&1; tail -1 $O/r02_smoke.log | cut -c1-120

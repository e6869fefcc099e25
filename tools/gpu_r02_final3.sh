#!/bin/sh
# round 2, last check of the tree as committed: GPU test suite and smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_final3.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final3.log
python -c "import __graft_entry__ as g; g.smoke()"

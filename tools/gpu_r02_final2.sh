#!/bin/sh
# round 2, final: GPU test suite and smoke() of the shipped build, then the round artefacts (bench lines, launch list, ncu summaries)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()"
R=r02 sh tools/gpu_round_artifacts.sh

#!/bin/sh
# round 2, call L2: division slow-path census (tools/ncu_div_slowpath.py) of the shading and trace kernels on boltsandgears (staged),
# complex.prc (single k_shade, persistent k_trace) and cornellbox_glassy
mkdir -p gpurun_out /tmp/reps
PRB_STAGED=1 timeout 400 ncu --section SourceCounters --clock-control none -k regex:"k_shade|k_trace" -s 60 -c 9 -o /tmp/reps/c4 -f python bench.py --scene c4 --no-cpu --no-extras --steps 1 --warmup 1 --spp 8 > gpurun_out/ncu_l2_c4.log 2>&1
PRB_STAGED=0 timeout 400 ncu --section SourceCounters --clock-control none -k regex:"k_shade|k_trace" -s 40 -c 3 -o /tmp/reps/c4c -f python bench.py --scene c4c --no-cpu --no-extras --steps 1 --warmup 1 --spp 4 > gpurun_out/ncu_l2_c4c.log 2>&1
PRB_STAGED=0 timeout 400 ncu --section SourceCounters --clock-control none -k regex:"k_shade|k_trace" -s 40 -c 2 -o /tmp/reps/c3 -f python bench.py --scene c3 --no-cpu --no-extras --steps 1 --warmup 1 --spp 8 > gpurun_out/ncu_l2_c3.log 2>&1
ls -la /tmp/reps
{
echo "== boltsandgears (staged)"; python tools/ncu_div_slowpath.py /tmp/reps/c4.ncu-rep "k_shade|k_trace"
echo "== complex.prc (single k_shade, persistent k_trace)"; python tools/ncu_div_slowpath.py /tmp/reps/c4c.ncu-rep "k_shade|k_trace"
echo "== cornellbox_glassy (single k_shade)"; python tools/ncu_div_slowpath.py /tmp/reps/c3.ncu-rep "k_shade|k_trace"
} > gpurun_out/r02_div_slowpath.txt 2>&1
cat gpurun_out/r02_div_slowpath.txt

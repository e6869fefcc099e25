#!/bin/sh
# Produces the per-round measurement artefacts on a B200 box (run through gpurun from the repository root); R = round tag.
#   gpurun_out/${R}_bench_<scene>.json    bench lines (numbers printed under ncu are never bench values)
#   gpurun_out/${R}_launches_c2.csv       ncu launch list of the default bench command (C2 leg)
#   gpurun_out/${R}_ncu_c2.txt            ncu --set full of k_trace_static / k_shade / k_regen on C2 in steady state: counters
#   gpurun_out/${R}_ncu_c4.txt            the same for the staged kernels on boltsandgears
#   gpurun_out/${R}_ncu_c5.txt            the same for the C5 ray-stream kernels (one launch each)
#   gpurun_out/${R}_hotspots_*.txt        samples / warp instructions / lanes per source line of the dominant kernels
# The .ncu-rep files stay on the box (gpurun brings back at most 64 MiB; three --import-source reports are 130 MB) except
# the C2 one when it fits; the summaries are made there with tools/ncu_summary.py / ncu_hotspots.py.
R=${R:-r02}
mkdir -p gpurun_out /tmp/reps
LIB=pearray_b200/libprb200.so
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py > gpurun_out/${R}_bench_c2.json 2> gpurun_out/${R}_bench_c2.err; cut -c1-400 gpurun_out/${R}_bench_c2.json
timeout 300 python bench.py --scene c5 > gpurun_out/${R}_bench_c5.json 2>/dev/null; cut -c1-200 gpurun_out/${R}_bench_c5.json
for sc in "c1 0" "c3 0" "c4 0" "c4c 256" "c0 0"; do set -- $sc  # complex.prc: 256 of its 4096 spp
  if [ "$2" = 0 ]; then spp=""; else spp="--spp $2"; fi
  timeout 900 python bench.py --scene $1 --no-cpu --steps 1 --warmup 1 $spp > gpurun_out/${R}_bench_$1.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/${R}_bench_$1.json')); print('$1', round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['stage_ms'], d.get('shading'))"
done
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/launch_run.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace|k_regen" -s 90 -c 3 -o /tmp/reps/${R}_c2 -f python bench.py --scene c2 --no-cpu --no-extras --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
PRB_STAGED=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace" -s 60 -c 9 -o /tmp/reps/${R}_c4 -f python bench.py --scene c4 --no-cpu --no-extras --steps 1 --warmup 1 --spp 8 > gpurun_out/ncu_c4.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o /tmp/reps/${R}_c5 -f python bench.py --scene c5 --no-cpu --steps 1 --warmup 1 --passes 1 > gpurun_out/ncu_c5.log 2>&1
ls -la /tmp/reps
for w in c2 c4 c5; do python tools/ncu_summary.py /tmp/reps/${R}_$w.ncu-rep --all > gpurun_out/${R}_ncu_$w.txt 2>&1; done
python tools/ncu_hotspots.py /tmp/reps/${R}_c2.ncu-rep k_shade $LIB 60 k_shadeILi128ELi1ELi2E > gpurun_out/${R}_hotspots_c2_shade.txt 2>&1
python tools/ncu_hotspots.py /tmp/reps/${R}_c2.ncu-rep k_trace_small $LIB 40 k_trace_small > gpurun_out/${R}_hotspots_c2_trace.txt 2>&1
python tools/ncu_hotspots.py /tmp/reps/${R}_c5.ncu-rep k_trace_closest $LIB 40 > gpurun_out/${R}_hotspots_c5_closest.txt 2>&1
python tools/ncu_hotspots.py /tmp/reps/${R}_c4.ncu-rep k_shade_geom $LIB 40 k_shade_geomILb0E > gpurun_out/${R}_hotspots_c4_geom.txt 2>&1
sz=$(stat -c %s /tmp/reps/${R}_c2.ncu-rep 2>/dev/null || echo 999999999)
if [ "$sz" -lt 45000000 ]; then cp /tmp/reps/${R}_c2.ncu-rep gpurun_out/; fi
du -sh gpurun_out

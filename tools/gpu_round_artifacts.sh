#!/bin/sh
# Produces the per-round measurement artefacts on a B200 box (run through gpurun from the repository root):
#   gpurun_out/r01_launches_c2_final.csv   ncu launch list of the default bench command
#   gpurun_out/r01_c5_10m_final.ncu-rep    ncu --set full of the C5 ray-stream kernels (one launch each)
#   gpurun_out/r01_c2_final.ncu-rep        ncu --set full of k_trace_static / k_shade on C2 in steady state
#   gpurun_out/bench_<scene>_final.json    bench lines (numbers printed under ncu are never bench values)
# Summaries for profiles/ are made from the .ncu-rep files with tools/ncu_summary.py / ncu_hotspots.py / ncu_sass_counts.py.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_c2_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launch_run.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/r01_c5_10m_final -f python bench.py --scene c5 --no-cpu --steps 1 --warmup 1 --passes 1 > gpurun_out/ncu_c5.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_trace" -s 60 -c 2 -o gpurun_out/r01_c2_final -f python bench.py --scene c2 --no-cpu --steps 1 --warmup 1 --spp 16 > gpurun_out/ncu_c2.log 2>&1
timeout 400 python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err; cat gpurun_out/bench_c2_final.json | cut -c1-300
timeout 300 python bench.py --scene c5 --no-cpu > gpurun_out/bench_c5_final.json 2>/dev/null; cat gpurun_out/bench_c5_final.json | cut -c1-200
for sc in "c1 0" "c3 0" "c4 0" "c4c 512"; do set -- $sc  # complex.prc: 512 of its 4096 spp (a full render is 140 s, the bench runs four)
  if [ "$2" = 0 ]; then spp=""; else spp="--spp $2"; fi
  timeout 900 python bench.py --scene $1 --no-cpu --steps 1 --warmup 1 $spp > gpurun_out/bench_$1_final.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/bench_$1_final.json')); print('$1', round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), d['stage_ms'])"
done

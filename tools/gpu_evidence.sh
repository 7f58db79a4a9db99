#!/bin/bash
# Evidence at HEAD (profiles/r02e_*, r02h_*, r02p_*; TAG=... selects the prefix): C2 bench line + launch list + one --set full capture of a whole frame,
# bench lines of C1 / C3 / C5, compute-sanitizer (memcheck, racecheck, synccheck, initcheck) on the smoke frame, the
# reference arm, the blend work counters, the whole GPU test suite.
mkdir -p gpurun_out
TAG=${TAG:-r02p}
TAG=$TAG WL=C2 bash tools/gpu_prof_all.sh
for wl in C1 C3 C5; do
  timeout 900 python bench.py --steps 100 --warmup 5 --workload $wl > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err
  tail -c 300 gpurun_out/${TAG}_bench_$wl.json; echo
done
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke\]|Error|hazard|SUMMARY" gpurun_out/${TAG}_sanitizer_$tool.log | cut -c1-200 | head -6
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_views.py > gpurun_out/${TAG}_sanitizer_views_$tool.log 2>&1
  echo "== $tool (renderer, lean state) rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|views\]|Error|hazard|SUMMARY" gpurun_out/${TAG}_sanitizer_views_$tool.log | cut -c1-200 | head -6
done
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_C2_reference_arm.json 2> gpurun_out/${TAG}_ref.err; tail -c 400 gpurun_out/${TAG}_bench_C2_reference_arm.json
timeout 600 python tools/blend_stats.py C2 > gpurun_out/${TAG}_blend_stats.txt 2>&1; timeout 600 python tools/blend_stats.py C5 >> gpurun_out/${TAG}_blend_stats.txt 2>&1; cat gpurun_out/${TAG}_blend_stats.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.txt; tail -3 gpurun_out/${TAG}_pytest_gpu.txt

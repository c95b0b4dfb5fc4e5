#!/bin/bash
# Per-build ncu evidence, in lock-step with the library: run ON THE GPU BOX (gpurun -- 'bash tools/ncu_capture.sh') after
# the bench has exited 0 without ncu.  Rebuilds libb200pt.so from the sources in this snapshot, captures the dominant
# kernel of every bench workload ONCE with `ncu --set full --import-source on` (after 3 warm-up launches) plus the launch
# list of the default bench, and records the hash of the kernel sources they were built from.  Back in the build container
# `python tools/ncu_counters.py` turns gpurun_out/ncu/*.ncu-rep into profiles/r02_*_summary.txt and
# profiles/ncu_counters.json; bench.py refuses that file when its hash differs from the sources it runs.
set -x
out=gpurun_out/ncu; mkdir -p $out; rm -f $out/captured.txt
make -C learn_path_tracing_b200/csrc -j8 > $out/build.log 2>&1 || { tail -20 $out/build.log; exit 1; }
python -c "import bench; print(bench.kernel_source_hash())" > $out/kernel_source_hash.txt
cat $out/kernel_source_hash.txt
# workload -> the (smaller-spp) form that is captured; same kernel, same scene, counters are rates
for pair in "10_final_720p_8192:10_final_720p_8192@1024" "8_refract_1080p:8_refract_1080p" "yoimiya_1080p:yoimiya_1080p" "zhongli_4k_4096:zhongli_4k" "intersect_10m:intersect_10m"; do
  name=${pair%%:*}; w=${pair##*:}; spp=""
  case "$w" in *@*) spp="--spp ${w##*@}"; w=${w%%@*};; esac   # the 8192-spp headline is captured at 1024 spp (same kernel, 8x shorter replay)
  [ -n "$NCU_ONLY" ] && [ "$NCU_ONLY" != "$name" ] && continue
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_paths_persist|k_trace_persist' \
      --launch-skip 3 -c 1 -f -o $out/r02_$name python bench.py --workload $w --only --steps 1 --no-cpu $spp > $out/ncu_$name.log 2>&1
  echo "$name captured_as=$w${spp:+@${spp##* }spp}" >> $out/captured.txt
  tail -2 $out/ncu_$name.log
done
if [ -z "$NCU_ONLY" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/r02_launches_default_bench.csv \
      python bench.py --steps 1 --no-cpu > $out/ncu_launches.log 2>&1
fi
ls -la $out

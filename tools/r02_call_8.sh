set -x
mkdir -p gpurun_out/r2h
out=gpurun_out/r2h/ab.txt; : > $out
V=$PWD/learn_path_tracing_b200/variants
for w in 10_final_720p 9_dof_720p 8_refract_1080p; do
  tools/sweep.sh $w "" default_sph >> $out
  PT_LIB_PATH=$V/libb200pt_nosph.so tools/sweep.sh $w "" nosph >> $out
  for sv in 8 16 20; do tools/sweep.sh $w "--serve-min $sv" sph_serve$sv >> $out; done
  for sm in 18 26; do tools/sweep.sh $w "--shade-min $sm" sph_shade$sm >> $out; done
  tools/sweep.sh $w "" default_sph >> $out
done
for v in brsh brsh24; do PT_LIB_PATH=$V/libb200pt_$v.so tools/sweep.sh intersect_10m "" $v >> $out; done
tools/sweep.sh intersect_10m "" default >> $out
cat $out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2h/gpu_tests.txt 2>&1; tail -n 5 gpurun_out/r2h/gpu_tests.txt

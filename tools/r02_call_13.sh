set -x
mkdir -p gpurun_out/r2m
out=gpurun_out/r2m/ab.txt; : > $out
# per-GPU shares of the strong-scaling jobs at N = 8 (and the full jobs), work units of 2 / 4 / 8 / 16 samples
for cfg in "yoimiya_1080p 64 2" "yoimiya_1080p 64 4" "yoimiya_1080p 64 8" "yoimiya_1080p 64 16" "yoimiya_1080p 256 4" "yoimiya_1080p 256 8" "yoimiya_1080p 256 16" "yoimiya_1080p 512 8" "yoimiya_1080p 512 16" \
           "8_refract_1080p 32 2" "8_refract_1080p 32 4" "8_refract_1080p 32 8" "8_refract_1080p 32 16" "8_refract_1080p 256 8" "8_refract_1080p 256 16" \
           "10_final_720p_8192 1024 8" "10_final_720p_8192 1024 16" "zhongli_4k 64 4" "zhongli_4k 64 8" "zhongli_4k 64 16" "zhongli_4k_4096 512 8" "zhongli_4k_4096 512 16"; do
  set -- $cfg
  PT_UNIT_SAMPLES=$3 tools/sweep.sh $1 "--spp $2" spp$2_unit$3 >> $out
done
cat $out

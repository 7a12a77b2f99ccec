set -x
mkdir -p gpurun_out
( for rep in 1 2; do for L in libccgpu.so libccgpu_r2y4.so libccgpu_r2y6.so; do
  export CCGPU_LIB=$PWD/channelcoding_b200/$L
  python tools/ab_fused.py --q 7 --t 10 --ebno 3 5
done; done
python tools/ab_fused.py --q 8 --t 18 --ebno 4 6
python tools/ab_fused.py --q 7 --t 10 --rows 127 --ebno 5 ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_r2y.txt

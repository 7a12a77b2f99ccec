set -x
mkdir -p gpurun_out
env | grep -i nccl
( for L in libccgpu.so libccgpu_lmb3.so libccgpu_lmb5.so; do
  export CCGPU_LIB=$PWD/channelcoding_b200/$L
  python tools/ab_fused.py --q 4 --t 2 --ebno 1 3 6 --variant MS --alpha 1.0
  python tools/ab_fused.py --q 4 --t 2 --ebno 1 3 --variant SPA --alpha 1.0
  python tools/ab_fused.py --q 4 --t 3 --ebno 3
done ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_lmb.txt

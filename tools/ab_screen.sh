set -x
mkdir -p gpurun_out
( for rep in 1 2; do for L in libccgpu.so libccgpu_ycol.so libccgpu_y30.so libccgpu_y32.so libccgpu_ycol30.so; do
  export CCGPU_LIB=$PWD/channelcoding_b200/$L
  python tools/ab_ms.py --q 6 --t 5 --ebno 4 --frames 4194304
done; done
for L in libccgpu.so libccgpu_ycol.so libccgpu_y30.so libccgpu_y32.so libccgpu_ycol30.so; do
  export CCGPU_LIB=$PWD/channelcoding_b200/$L
  python tools/ab_ms.py --q 6 --t 3 --ebno 4 --frames 4194304
  python tools/ab_ms.py --q 6 --t 4 --ebno 4 --frames 4194304
  python tools/ab_ms.py --q 5 --t 4 --ebno 4 --frames 8388608
  python tools/ab_ms.py --q 6 --t 5 --ebno 7 --frames 8388608
done ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_ycol.txt

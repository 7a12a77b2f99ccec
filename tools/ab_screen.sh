set -x
mkdir -p gpurun_out
for L in libccgpu_base.so libccgpu.so libccgpu_qmb7.so libccgpu_qmb6.so; do
  export CCGPU_LIB=$PWD/channelcoding_b200/$L
  python tools/ab_fused.py --q 7 --t 10 --rows 127 --ebno 3 5 7 --variant NMS_Q
  python tools/ab_fused.py --q 7 --t 10 --rows 127 --ebno 5
done 2>&1 | tee gpurun_out/ab_screen3.txt
export CCGPU_LIB=$PWD/channelcoding_b200/libccgpu.so
python tools/ab_fused.py --q 6 --t 5 --ebno 4 5 5.5 --quick 1 2>&1 | tee -a gpurun_out/ab_screen3.txt
python tools/ab_fused.py --q 6 --t 5 --ebno 4 5 5.5 --quick 0 2>&1 | tee -a gpurun_out/ab_screen3.txt
python tools/ab_fused.py --q 7 --t 10 --ebno 5 6 7 --quick 1 2>&1 | tee -a gpurun_out/ab_screen3.txt
python tools/ab_fused.py --q 7 --t 10 --ebno 5 6 7 --quick 0 2>&1 | tee -a gpurun_out/ab_screen3.txt

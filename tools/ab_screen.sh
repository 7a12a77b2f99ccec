set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_lane.py -x -q -m gpu 2>&1 | tail -15
export CCGPU_LIB=$PWD/channelcoding_b200/libccgpu.so
( for V in "MS --alpha 1.0" "NMS" "SPA --alpha 1.0"; do
  CCGPU_LANE=0 python tools/ab_fused.py --q 4 --t 2 --ebno 1 3 6 9 --variant $V --lane 0
  python tools/ab_fused.py --q 4 --t 2 --ebno 1 3 6 9 --variant $V --lane 1
done
python tools/ab_fused.py --q 4 --t 3 --ebno 1 3 6 --lane 0
python tools/ab_fused.py --q 4 --t 3 --ebno 1 3 6 --lane 1
python tools/ab_ms.py --q 4 --t 2 --ebno 3 --variant MS --alpha 1.0 --frames 16777216
python tools/ab_fused.py --q 7 --t 10 --rows 127 --ebno 5 --quick 0
python tools/ab_fused.py --q 7 --t 10 --rows 127 --ebno 5 --quick 1
) 2>&1 | tee gpurun_out/ab_lane1.txt

set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_lane.py -x -q -m gpu 2>&1 | tail -5
( python tools/ab_fused.py --q 4 --t 2 --ebno 1 3 6 9 --variant MS --alpha 1.0
python tools/ab_fused.py --q 4 --t 2 --ebno 1 3 6 --variant NMS
python tools/ab_fused.py --q 4 --t 2 --ebno 3 --variant OMS --alpha 1.0
python tools/ab_fused.py --q 4 --t 3 --ebno 1 3 6
) 2>&1 | tee gpurun_out/ab_lane2.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -5

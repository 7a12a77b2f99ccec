set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or every_compiled or screening or shortcut" 2>&1 | tail -3
python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
( for LN in 0 1; do
  python tools/ab_fused.py --q 5 --t 3 --ebno 2 4 6 8 --lane $LN
  python tools/ab_fused.py --q 5 --t 1 --ebno 4 6 --lane $LN
  python tools/ab_fused.py --q 5 --t 2 --ebno 4 6 --lane $LN
  python tools/ab_fused.py --q 5 --t 4 --ebno 2 4 --lane $LN
  python tools/ab_fused.py --q 5 --t 3 --ebno 4 --variant SPA --alpha 1.0 --lane $LN
done ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_lane31.txt

set -x
mkdir -p gpurun_out
( CCGPU_LIB=$PWD/channelcoding_b200/libccgpu_base.so python tools/ab_k1.py
python tools/ab_k1.py ) 2>&1 | grep -v "^+" | tee gpurun_out/ab_k1.txt
python tools/sweep.py --q 8 --t 18 --variant NMS --alpha 0.8 --ebno-from 0 --ebno-to 11.5 --ebno-step 0.5 --max-frames 1e10 > gpurun_out/waterfall_255_131_n1_r2b.jsonl 2> gpurun_out/waterfall_r2b.err
tail -4 gpurun_out/waterfall_255_131_n1_r2b.jsonl | cut -c1-330
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "channel" 2>&1 | tail -3

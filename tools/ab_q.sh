for lib in libccgpu.so libccgpu_q0.so; do
  for eb in 4 6 7 8; do CCGPU_LIB=$PWD/channelcoding_b200/$lib python tools/ab_ms.py --variant NMS_Q --ebno $eb --frames 8388608 --reps 3; done
  CCGPU_LIB=$PWD/channelcoding_b200/$lib python tools/ab_ms.py --q 4 --t 2 --variant MS_Q --ebno 3 --frames 33554432 --reps 3
  CCGPU_LIB=$PWD/channelcoding_b200/$lib python tools/ab_ms.py --q 4 --t 2 --variant MS_Q --ebno 7 --frames 33554432 --reps 3
  CCGPU_LIB=$PWD/channelcoding_b200/$lib python tools/ab_ms.py --q 7 --t 10 --variant NMS_Q --ebno 7 --frames 4194304 --reps 3
  CCGPU_LIB=$PWD/channelcoding_b200/$lib python tools/ab_ms.py --q 8 --t 18 --variant NMS_Q --ebno 8.5 --frames 1048576 --reps 3
done

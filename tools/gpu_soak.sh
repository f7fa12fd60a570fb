#!/bin/bash
# soak: the seeded option-space fuzz of tests/test_gpu_parity.py over many more cases than the suite runs (bit-exact against the oracle)
OUT=gpurun_out/${1:-soak}
mkdir -p $OUT
CHB_FUZZ_BASE=${2:-1000} CHB_FUZZ_CASES=${3:-3000} timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "random_option_fuzz" -p no:cacheprovider > $OUT/soak.txt 2>&1
tail -5 $OUT/soak.txt
CHB_FUZZ_BASE=${2:-1000} CHB_VIDEO_FUZZ_CASES=${4:-1500} timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "video_run_fuzz" -p no:cacheprovider > $OUT/soak_video.txt 2>&1
tail -5 $OUT/soak_video.txt
CHB_FUZZ_BASE=${2:-1000} CHB_SEQ_FUZZ_CASES=${5:-300} timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "call_sequences" -p no:cacheprovider > $OUT/soak_seq.txt 2>&1
tail -5 $OUT/soak_seq.txt
CHB_FUZZ_BASE=${2:-1000} CHB_SIMPLE_FUZZ_CASES=${6:-3000} timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "simple_fuzz" -p no:cacheprovider > $OUT/soak_simple.txt 2>&1
tail -5 $OUT/soak_simple.txt

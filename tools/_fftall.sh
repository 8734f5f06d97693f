export KBENCH_TOTAL=$((1<<28))
python tools/kbench.py fft 128 256 512 1024 2048 4096 8192 16384 4410 4800 9600 19200 2>&1 | grep "fft n"
JSDR_FFT_ALT=9600:b python tools/kbench.py fft 9600 2>&1 | grep "fft n"
JSDR_FFT_ALT=19200:c python tools/kbench.py fft 19200 2>&1 | grep "fft n"
JSDR_FFT_ALT=16384:b python tools/kbench.py fft 16384 2>&1 | grep "fft n"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fft 2>&1 | tail -2

export KBENCH_TOTAL=$((1<<28))
echo "== defaults"; python tools/kbench.py fft 9600 16384 19200 2>&1 | grep -v "^$"
for a in 19200:a 19200:b 19200:c 19200:d; do echo "== JSDR_FFT_ALT=$a"; JSDR_FFT_ALT=$a python tools/kbench.py fft 19200 2>&1 | tail -2; done
for a in 9600:a 9600:b 9600:c; do echo "== JSDR_FFT_ALT=$a"; JSDR_FFT_ALT=$a python tools/kbench.py fft 9600 2>&1 | tail -2; done
for a in 16384:a 16384:b; do echo "== JSDR_FFT_ALT=$a"; JSDR_FFT_ALT=$a python tools/kbench.py fft 16384 2>&1 | tail -2; done

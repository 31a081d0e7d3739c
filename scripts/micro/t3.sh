t() { timeout 120 python scripts/run_one.py --D 192 --reps 4 | tail -2 | head -1 | tr "," "\n" | grep scanline | grep -v "#n" | tr "\n" " "; echo; }
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scanline3.py tests/test_gpu_hsi.py tests/test_gpu_determinism.py tests/test_gpu_mindisp.py -x -q 2>&1 | tail -3
echo default; t; t; t
for D in 48 192; do timeout 300 python scripts/stress_mixed.py $D 30 2>&1 | tail -2; done

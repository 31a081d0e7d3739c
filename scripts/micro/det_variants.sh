for v in scan_dep scan_fence scan_late; do
  echo "== $v"; TSM_LIB=$PWD/scripts/micro/libs/$v.so timeout 300 python -m pytest tests/test_gpu_determinism.py -x -q -k next_to 2>&1 | grep -E "differ|passed|failed" | head -3
done

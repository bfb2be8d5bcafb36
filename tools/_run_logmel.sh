timeout 300 python -m pytest tests/test_gpu_logmel.py -x -q 2>&1 | tail -4
for c in "" 5 6; do
  echo "== cluster '$c'"; WSB_LOGMEL_CLUSTER=$c timeout 120 python tools/logmel_one.py 2>&1 | tail -1
  WSB_LOGMEL_CLUSTER=$c timeout 120 python tools/logmel_one.py 16000 0.01 3600 2>&1 | tail -1
done
timeout 120 python tools/logmel_one.py 32000 0.0025 600 2>&1 | tail -1
timeout 120 python tools/logmel_one.py 44100 0.0025 600 2>&1 | tail -1

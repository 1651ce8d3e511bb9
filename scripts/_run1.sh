set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor_core_updates" 2>&1 | tail -15

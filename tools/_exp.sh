python -m pytest tests/test_gpu_process_image.py -m gpu -x -q -k "sweep" 2>&1 | tail -12

python -m pytest tests/test_gpu_process_image.py -m gpu -x -q 2>&1 | tail -15

python -m pytest tests/test_gpu_motion.py -m gpu -x -q -k "sweep or corners" 2>&1 | tail -12

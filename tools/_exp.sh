python -m pytest tests/test_gpu_cli.py -m gpu -x -q 2>&1 | tail -3
python tools/cli_throughput.py --frames 120 --format png
python tools/cli_throughput.py --frames 120 --format pgm

L=pixeltovoxelprojector_b200/libpvp_b200.so
cp $L /tmp/real.so; cp exp_build/oldpark.so $L
python -m pytest tests/test_gpu_motion.py -m gpu -x -q -k "low_corners" 2>&1 | tail -6
cp /tmp/real.so $L

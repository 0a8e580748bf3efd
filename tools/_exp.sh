source tools/gpu_exp.sh
L=pixeltovoxelprojector_b200/libpvp_b200.so
cp $L /tmp/real.so
for e in real t256 t64 real t256; do
  if [ $e = real ]; then cp /tmp/real.so $L; else cp exp_build/$e.so $L; fi
  run ${e}_r80_512 --workload 1080p_dense_512 --opt lean_regs=80
done
cp /tmp/real.so $L

for k in 0 1 3; do echo "POLY_PAIRS=$k"; XB_LIB=$PWD/matrix-factorization-torch_b200/libxfmr_b200_p$k.so python tests/quick_probe.py 2>&1 | grep -E "Loss:"; done
echo "POLY_PAIRS=2 (default)"; python tests/quick_probe.py 2>&1 | grep -E "Loss:"

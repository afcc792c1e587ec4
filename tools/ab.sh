python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "persist or multi or gemv" 2>&1 | tail -3
export PTIME_VARIANTS="1 0 0;0 0 0"
echo "### new"; python tools/ptime.py 4096 4096 4096 11008 11008 4096 8192 8192 | grep "persist \|=="
echo "### nopin(prev)"; XBIT_B200_LIB=$PWD/_ab/exp/lib_nopin.so python tools/ptime.py 4096 4096 4096 11008 11008 4096 8192 8192 | grep "persist fine"

for v in CLAST PRELOOP; do echo "### $v"; XBIT_B200_LIB=$PWD/_ab/exp/lib_dev$v.so TRACE_FAMILY=5 python tools/trace.py 8192 28672; done

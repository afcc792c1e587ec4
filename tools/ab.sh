python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "generic or gemv_all_bits or ragged or w8 or every_width or bits" 2>&1 | tail -5
python tools/pw8.py --bits 3 4096 4096 4096 11008 11008 4096
python tools/pw8.py --bits 8 4096 11008
python tools/pw8.py --bits 6 --m 4 4096 11008

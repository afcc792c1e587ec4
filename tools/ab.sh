python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "w8 or bf16_native" 2>&1 | tail -12
python tools/pw8.py --bits 2
python tools/pw8.py --bits 8 4096 11008

export PTIME_VARIANTS="1 0 0;0 0 0"
for m in 2 4 8; do python tools/ptime.py --m $m 4096 4096 4096 11008 8192 8192 | grep "persist fine\|==\|cluster\|AUTO"; done

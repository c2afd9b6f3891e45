DENSE=1 python tools/phase_timing.py 2>&1 | head -3

timeout 900 python tools/_config3.py 2>&1 | tail -8

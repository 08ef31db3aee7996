timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pinned or separately or batch" 2>&1 | tail -4
timeout 300 python tools/commit_stages.py 2>&1 | tail -8

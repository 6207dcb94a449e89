# adaptive schemes on a partitioned body: world size 1 (one GPU) through pytest
timeout -k 10 900 python -m pytest tests/test_gpu_dist_single.py -m gpu -x -q 2>&1 | tail -30

# W-method attempt graphs: new test + the W-method parity tests through the graph path + latency A/B
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "w_method or ros2 or row3" 2>&1 | tail -5
timeout -k 10 900 python scripts/wmethod_graph_latency.py --json gpurun_out/r02_wmethod_graph.json 2>&1 | tail -20

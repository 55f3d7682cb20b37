"""BASELINE metric 2 alone: bench.solve_full_size (sweep -> Newton certification on the 4097^2
lattice) as one JSON line.   python tools/solve4097.py [N] [--no-cpu-baseline] [--solve-policy-from-g]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

n = [a for a in sys.argv[1:] if a.isdigit()]
sys.argv = [sys.argv[0]] + [a for a in sys.argv[1:] if not a.isdigit()] + (["--solve-full", n[0]] if n else [])
args = bench.parse()
print(json.dumps(bench.solve_full_size(args, bench.two_cones)))

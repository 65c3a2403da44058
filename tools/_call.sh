python tools/explain_bench.py 2>&1 | tail -5

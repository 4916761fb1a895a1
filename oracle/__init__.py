"""TEST INFRASTRUCTURE ONLY -- never imported by the product package `redgnn_b200`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker or the reported CPU baseline.
"""

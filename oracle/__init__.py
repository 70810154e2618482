"""TEST INFRASTRUCTURE ONLY.

CPU fp32 restatement of the cc-ai/MUNIT hot path (networks.py / trainer.py /
extraadam.py).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import this package.  The product
(`munit_b200`) never imports it and fails loudly without its CUDA extension.
"""

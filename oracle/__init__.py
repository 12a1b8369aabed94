"""ORACLE — test infrastructure only.

CPU restatement of the reference's ray-rendering hot path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import this package, and only as the checker or the timed CPU baseline — never as a product path.
"""

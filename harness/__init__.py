"""Pure-Python test/bench harness helpers: synthetic signals and size/shard arithmetic.

Nothing in this package loads native code, so `bench.py --impl reference|cpu` can use it without mapping the
product library (the driver records which .so files each arm loads).
"""

"""Multi-GPU arm of bench.py (filled in once the merge path is built)."""


def run(args, *a):
    raise SystemExit("multi-GPU bench not built yet")

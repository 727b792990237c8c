"""Reference arm of bench.py: the reference's own PyTorch code (baseline/_ref, installed by baseline/make_ref.py)."""

"""Bare-name shim: `import MPGCN` (as the reference's Model_Trainer.py:5 does) resolves to the
B200 engine's drop-in classes when this repository precedes the reference on sys.path."""
from mpgcn_b200.MPGCN import BDGCN, MPGCN  # noqa: F401

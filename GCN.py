"""Bare-name shim: `import GCN` (as the reference's Model_Trainer.py:5 does) resolves to the B200 engine's support-matrix
builder when this repository precedes the reference on sys.path."""
from mpgcn_b200.GCN import Adj_Processor  # noqa: F401

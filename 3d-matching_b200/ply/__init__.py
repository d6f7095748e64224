"""`from ply import Ply` (src/main.py:15, benchmark_ransac.py:21) resolves here: the Ply container on the B200 engine."""
from ply.ply import Ply as Ply

__all__ = ("Ply",)

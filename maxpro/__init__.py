"""`import maxpro` -- the reference's Python module name (pyo3, src/lib.rs:166-176) bound to the
B200 engine, so the reference's own scripts (python/comparison_r.py, comparisons.py, figures.py)
run unmodified with this repository on sys.path.  Everything lives in maxpro_b200.maxpro."""
from maxpro_b200.maxpro import (  # noqa: F401
    anneal_lhd,
    build_lhd,
    generate_lhd,
    maximin_criterion,
    maxpro_criterion,
    order_design,
)

__all__ = ["anneal_lhd", "build_lhd", "generate_lhd", "maximin_criterion", "maxpro_criterion", "order_design"]

"""Fixture tooling: the `.rds` reader used by tests/golden/make_golden.py. The codec itself lives in the package
(`magmaclust_b200/rds_io.py`, reader + writer); this module only keeps the old import path working."""
from magmaclust_b200.rds_io import RList, RStrings, RVector, read_rds  # noqa: F401

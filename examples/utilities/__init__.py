"""Shared helpers of the example scripts (the reference's ``examples/utilities/__init__.py:7-13`` imports modules that
do not exist — SURVEY §8-D17; here it simply re-exports :mod:`examples.common`)."""
from examples.common import *  # noqa: F401,F403

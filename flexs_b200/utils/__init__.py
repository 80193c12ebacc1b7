"""Utilities; ``sequence_utils`` holds the alphabets and sequence manipulation helpers."""
from flexs_b200.utils import sequence_utils  # noqa: F401

"""Synthetic stand-in for the reference's missing ``data/dataset.py``."""

"""Settings (brie/settings.py:1-6)."""

verbosity = 3
"""Verbosity level (0=errors, 1=warnings, 2=info, 3=hints)"""

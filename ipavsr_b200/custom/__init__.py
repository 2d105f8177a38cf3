"""Mirrors of the reference's `custom/` package (layers, objectives, updates, nonlinearities)."""

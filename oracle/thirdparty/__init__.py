"""Restated third-party algorithms (e3nn, torch-scatter, torch-cluster). Test infrastructure only."""
